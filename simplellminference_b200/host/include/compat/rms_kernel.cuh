// compat/rms_kernel.cuh — forwarding header: code written against the reference includes "rms_kernel.cuh"; here it lives in sllm/kernel.h.
#pragma once
#include "../sllm/kernel.h"
