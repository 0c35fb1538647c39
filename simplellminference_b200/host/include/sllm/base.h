// sllm/base.h — error convention and device/memcpy enums of the reference (include/base/base.h:6-24).
// Same names and meaning: LOG(msg) prints "file: <f> line: <n> - <msg>" and terminates the process.
#pragma once
#include <cstdlib>
#include <iostream>
#include <string>

namespace sllm_host {
[[noreturn]] inline void fatal(const std::string& message, const char* file, int line) {
    std::cout << "file: " << file << " line: " << line << " - " << message << std::endl;
    std::exit(EXIT_FAILURE);
}
}  // namespace sllm_host
#define LOG(message) ::sllm_host::fatal((message), __FILE__, __LINE__)

namespace base {
enum class MemcpyKind { kMemcpyCPU2CPU = 0, kMemcpyCPU2CUDA = 1, kMemcpyCUDA2CPU = 2, kMemcpyCUDA2CUDA = 3 };
enum class DeviceType { kDeviceUnknown = 0, kDeviceCPU = 1, kDeviceCUDA = 2 };
// extension: element type of a tensor (the reference is fp32-only: `#define DataTypeSize 4`, alloc.h:12)
enum class DataType { kFp32 = 0, kBf16 = 1, kInt8 = 2, kInt32 = 3 };
inline size_t data_type_size(DataType t) { return t == DataType::kBf16 ? 2 : t == DataType::kInt8 ? 1 : 4; }
}  // namespace base
