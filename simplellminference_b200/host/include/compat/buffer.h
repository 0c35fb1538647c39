// compat/buffer.h — forwarding header: code written against the reference includes "buffer.h"; here it lives in sllm/memory.h.
#pragma once
#include "../sllm/memory.h"
