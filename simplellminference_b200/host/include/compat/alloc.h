// compat/alloc.h — forwarding header: code written against the reference includes "alloc.h"; here it lives in sllm/memory.h.
#pragma once
#include "../sllm/memory.h"
