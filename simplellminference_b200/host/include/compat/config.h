// compat/config.h — forwarding header: code written against the reference includes "config.h"; here it lives in sllm/model.h.
#pragma once
#include "../sllm/model.h"
