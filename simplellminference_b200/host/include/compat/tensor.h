// compat/tensor.h — forwarding header: code written against the reference includes "tensor.h"; here it lives in sllm/memory.h.
#pragma once
#include "../sllm/memory.h"
