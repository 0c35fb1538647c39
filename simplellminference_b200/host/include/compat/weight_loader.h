// compat/weight_loader.h — forwarding header: code written against the reference includes "weight_loader.h"; here it lives in sllm/model.h.
#pragma once
#include "../sllm/model.h"
