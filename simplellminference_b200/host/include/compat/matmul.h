// compat/matmul.h — forwarding header: code written against the reference includes "matmul.h"; here it lives in sllm/op.h.
#pragma once
#include "../sllm/op.h"
