// compat/emb_kernel.cuh — forwarding header: code written against the reference includes "emb_kernel.cuh"; here it lives in sllm/kernel.h.
#pragma once
#include "../sllm/kernel.h"
