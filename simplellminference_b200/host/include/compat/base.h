// compat/base.h — forwarding header: code written against the reference includes "base.h"; here it lives in sllm/base.h.
#pragma once
#include "../sllm/base.h"
