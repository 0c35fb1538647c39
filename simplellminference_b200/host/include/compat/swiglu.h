// compat/swiglu.h — forwarding header: code written against the reference includes "swiglu.h"; here it lives in sllm/op.h.
#pragma once
#include "../sllm/op.h"
