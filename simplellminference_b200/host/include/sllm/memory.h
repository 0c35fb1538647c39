// sllm/memory.h — mem::DeviceAllocator / Buffer / Tensor with the reference's contracts
// (include/memory/alloc.h:32-137, buffer.h:18-55, tensor.h:10-127), re-implemented:
//  * CUDADeviceAllocator is a slab pool with address-ordered coalescing (the reference's pool mixes remnants
//    between its two size classes and never returns slabs); allocate/release/memcpy/
//    memset_zero keep their signatures and error behaviour (std::bad_alloc on device OOM after a cache flush);
//  * Tensor carries an element type (fp32 default) instead of the hard-wired 4 bytes; sizes are 64-bit;
//  * Buffer/Tensor ownership rules are the reference's: a Tensor is a value sharing a shared_ptr<Buffer>;
//    a Buffer frees through its allocator unless it wraps external memory.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

#include "base.h"

namespace mem {

inline size_t round_size(size_t size) { return size < 512 ? 512 : (size + 511) / 512 * 512; }

class DeviceAllocator {
public:
    explicit DeviceAllocator(base::DeviceType device_type) : device_type_(device_type) {}
    virtual ~DeviceAllocator() = default;
    virtual base::DeviceType device_type() { return device_type_; }
    virtual void release(void* ptr) const = 0;
    virtual void* allocate(size_t byte_size) const = 0;
    void memcpy(const void* src_ptr, void* dst_ptr, size_t byte_size, base::MemcpyKind memcpy_kind) const;
    virtual void memset_zero(void* ptr, size_t byte_size);

private:
    base::DeviceType device_type_ = base::DeviceType::kDeviceUnknown;
};

class CPUDeviceAllocator : public DeviceAllocator {
public:
    CPUDeviceAllocator() : DeviceAllocator(base::DeviceType::kDeviceCPU) {}
    void* allocate(size_t byte_size) const override;
    void release(void* ptr) const override;
};

// Device memory pool: cudaMalloc'ed slabs carved into blocks; free blocks are kept in one address-ordered map
// per slab list and merged with both neighbours on release; best fit by size.
class CUDADeviceAllocator : public DeviceAllocator {
public:
    CUDADeviceAllocator() : DeviceAllocator(base::DeviceType::kDeviceCUDA) {}
    ~CUDADeviceAllocator() override;
    void* allocate(size_t byte_size) const override;
    void release(void* ptr) const override;
    size_t bytes_reserved() const;   // slabs held
    size_t bytes_in_use() const;     // handed out
    void release_cached_memory() const;   // give fully-free slabs back to the driver

private:
    struct Block { size_t size; bool used; size_t slab; };
    mutable std::mutex mu_;
    mutable std::map<uintptr_t, Block> blocks_;            // every block (used or free), by address
    mutable std::multimap<size_t, uintptr_t> free_;        // free blocks by size
    mutable std::vector<std::pair<void*, size_t>> slabs_;
    mutable size_t in_use_ = 0;
    void drop_free(uintptr_t addr, size_t size) const;
};

class CPUDeviceAllocatorFactory {
public:
    static std::shared_ptr<CPUDeviceAllocator> get_instance();
};
class CUDADeviceAllocatorFactory {
public:
    static std::shared_ptr<CUDADeviceAllocator> get_instance();
};

class Buffer {
public:
    Buffer() = default;
    explicit Buffer(size_t byte_size, std::shared_ptr<DeviceAllocator> allocator = nullptr, void* ptr = nullptr,
                    bool use_external = false);
    virtual ~Buffer();
    Buffer(const Buffer&) = delete;
    Buffer& operator=(const Buffer&) = delete;

    bool allocate();
    void copy_from(const Buffer& buffer) const;
    void copy_from(const Buffer* buffer) const;
    void* ptr() { return ptr_; }
    const void* ptr() const { return ptr_; }
    size_t byte_size() const { return byte_size_; }
    std::shared_ptr<DeviceAllocator> allocator() const { return allocator_; }
    base::DeviceType device_type() const { return device_type_; }
    void set_device_type(base::DeviceType device_type) { device_type_ = device_type; }
    bool is_external() const { return use_external_; }

private:
    size_t byte_size_ = 0;
    void* ptr_ = nullptr;
    bool use_external_ = false;
    base::DeviceType device_type_ = base::DeviceType::kDeviceUnknown;
    std::shared_ptr<DeviceAllocator> allocator_;
};

class Tensor {
public:
    Tensor() = default;
    explicit Tensor(std::vector<int32_t> dims, bool need_alloc = false, std::shared_ptr<DeviceAllocator> alloc = nullptr,
                    void* ptr = nullptr, base::DataType dtype = base::DataType::kFp32);

    void to_cpu();
    void to_cuda();
    bool is_empty() const;
    void init_buffer(std::shared_ptr<DeviceAllocator> alloc, bool need_alloc, void* ptr);
    void reshape(const std::vector<int32_t>& dims);
    std::shared_ptr<Buffer> get_buffer() const { return buffer_; }
    size_t size() const { return size_; }
    size_t byte_size() const { return size_ * base::data_type_size(dtype_); }
    int32_t dims_size() const { return static_cast<int32_t>(dims_.size()); }
    int32_t get_dim(int32_t idx) const;
    const std::vector<int32_t>& dims() const { return dims_; }
    std::vector<size_t> strides() const;
    bool assign(std::shared_ptr<Buffer> buffer);
    void reset(const std::vector<int32_t>& dims);
    void set_device_type(base::DeviceType device_type) const;
    base::DeviceType device_type() const;
    bool allocate(std::shared_ptr<DeviceAllocator> allocator, bool need_realloc = false);
    Tensor clone() const;
    base::DataType data_type() const { return dtype_; }

    template <typename T> T* ptr() { return buffer_ ? reinterpret_cast<T*>(buffer_->ptr()) : nullptr; }
    template <typename T> const T* ptr() const { return buffer_ ? reinterpret_cast<const T*>(buffer_->ptr()) : nullptr; }
    template <typename T> T* ptr(int64_t index) {
        if (!buffer_ || !buffer_->ptr()) LOG("ERROR Get Ptr!");
        return reinterpret_cast<T*>(buffer_->ptr()) + index;
    }
    template <typename T> const T* ptr(int64_t index) const {
        if (!buffer_ || !buffer_->ptr()) LOG("ERROR Get Ptr!");
        return reinterpret_cast<const T*>(buffer_->ptr()) + index;
    }
    // host-side element access (CPU tensors only, like the reference's use of index<>() on token/position)
    template <typename T> T& index(int64_t offset) {
        if (offset < 0 || offset >= static_cast<int64_t>(size_)) LOG("ERROR Index!");
        return *(reinterpret_cast<T*>(buffer_->ptr()) + offset);
    }
    template <typename T> const T& index(int64_t offset) const {
        if (offset < 0 || offset >= static_cast<int64_t>(size_)) LOG("ERROR Index!");
        return *(reinterpret_cast<const T*>(buffer_->ptr()) + offset);
    }

private:
    size_t size_ = 0;
    std::vector<int32_t> dims_;
    std::shared_ptr<Buffer> buffer_;
    base::DataType dtype_ = base::DataType::kFp32;
};

// non-owning {dim} views of cache rows (layer_idx, pos) — reference tensor.cpp:199-212, with 64-bit offsets
std::pair<Tensor, Tensor> slice_KV_cache(int32_t layer_idx, int32_t pos, int32_t max_seq_len, int32_t dim,
                                         const Tensor& key_cache, const Tensor& value_cache);

}  // namespace mem
