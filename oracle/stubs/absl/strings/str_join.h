// Empty stub: op/encode.h includes this third-party header but uses nothing from it.
#pragma once
