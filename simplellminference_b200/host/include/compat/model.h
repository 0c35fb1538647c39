// compat/model.h — forwarding header: code written against the reference includes "model.h"; here it lives in sllm/model.h.
#pragma once
#include "../sllm/model.h"
