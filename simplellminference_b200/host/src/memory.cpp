// host/src/memory.cpp — allocators, Buffer, Tensor (contracts: reference include/memory/*.h, source/memory/*.cpp).
#include <cuda_runtime_api.h>

#include <new>
#include <numeric>

#include "sllm/memory.h"

namespace mem {

// ------------------------------------------------------------------------------------ DeviceAllocator ----
void DeviceAllocator::memcpy(const void* src_ptr, void* dst_ptr, size_t byte_size, base::MemcpyKind kind) const {
    if (!src_ptr || !dst_ptr) LOG(" ERROR! Ptr is empty! ");
    if (kind == base::MemcpyKind::kMemcpyCPU2CPU) {
        std::memcpy(dst_ptr, src_ptr, byte_size);
        return;
    }
    const cudaMemcpyKind k = kind == base::MemcpyKind::kMemcpyCPU2CUDA   ? cudaMemcpyHostToDevice
                             : kind == base::MemcpyKind::kMemcpyCUDA2CPU ? cudaMemcpyDeviceToHost
                                                                         : cudaMemcpyDeviceToDevice;
    const cudaError_t e = cudaMemcpy(dst_ptr, src_ptr, byte_size, k);   // the reference ignores CUDA errors; we do not
    if (e != cudaSuccess) LOG(std::string("cudaMemcpy failed: ") + cudaGetErrorString(e));
}

void DeviceAllocator::memset_zero(void* ptr, size_t byte_size) {
    if (!ptr) LOG(" ERROR! Ptr is Empty! ");
    if (device_type_ == base::DeviceType::kDeviceUnknown) LOG(" ERROR! Device Type Unknown! ");
    if (device_type_ == base::DeviceType::kDeviceCPU) {
        std::memset(ptr, 0, byte_size);
    } else if (cudaMemset(ptr, 0, byte_size) != cudaSuccess) {
        LOG("cudaMemset failed");
    }
}

void* CPUDeviceAllocator::allocate(size_t byte_size) const { return byte_size ? std::malloc(byte_size) : nullptr; }
void CPUDeviceAllocator::release(void* ptr) const { std::free(ptr); }

// --------------------------------------------------------------------------------- CUDADeviceAllocator ----
namespace {
constexpr size_t kSlabSmall = 8u << 20;    // requests <= 1 MiB share 8 MiB slabs
constexpr size_t kSmallLimit = 1u << 20;
constexpr size_t kSlabAlign = 2u << 20;    // larger requests get their own slab, rounded to 2 MiB
}  // namespace

CUDADeviceAllocator::~CUDADeviceAllocator() {
    for (auto& s : slabs_)
        if (s.first) cudaFree(s.first);
}

void CUDADeviceAllocator::drop_free(uintptr_t addr, size_t size) const {
    auto range = free_.equal_range(size);
    for (auto it = range.first; it != range.second; ++it)
        if (it->second == addr) { free_.erase(it); return; }
}

void* CUDADeviceAllocator::allocate(size_t byte_size) const {
    if (!byte_size) return nullptr;
    const size_t size = round_size(byte_size);
    std::lock_guard<std::mutex> lock(mu_);
    auto fit = free_.lower_bound(size);   // best fit
    if (fit == free_.end()) {
        const size_t slab_bytes = size <= kSmallLimit ? kSlabSmall : (size + kSlabAlign - 1) / kSlabAlign * kSlabAlign;
        void* p = nullptr;
        if (cudaMalloc(&p, slab_bytes) != cudaSuccess) {
            cudaGetLastError();
            mu_.unlock();
            release_cached_memory();   // same recovery order as the reference (alloc.cpp:118-131)
            mu_.lock();
            if (cudaMalloc(&p, slab_bytes) != cudaSuccess) { cudaGetLastError(); throw std::bad_alloc(); }
        }
        slabs_.emplace_back(p, slab_bytes);
        blocks_[reinterpret_cast<uintptr_t>(p)] = Block{slab_bytes, false, slabs_.size() - 1};
        fit = free_.emplace(slab_bytes, reinterpret_cast<uintptr_t>(p));
    }
    const uintptr_t addr = fit->second;
    free_.erase(fit);
    Block& b = blocks_[addr];
    if (b.size - size >= 512) {   // split: the remainder stays free
        blocks_[addr + size] = Block{b.size - size, false, b.slab};
        free_.emplace(b.size - size, addr + size);
        b.size = size;
    }
    b.used = true;
    in_use_ += b.size;
    return reinterpret_cast<void*>(addr);
}

void CUDADeviceAllocator::release(void* ptr) const {
    if (!ptr) return;
    std::lock_guard<std::mutex> lock(mu_);
    auto it = blocks_.find(reinterpret_cast<uintptr_t>(ptr));
    if (it == blocks_.end() || !it->second.used) return;   // unknown pointer: ignored, like the reference
    it->second.used = false;
    in_use_ -= it->second.size;
    auto next = std::next(it);   // merge with the following free block of the same slab
    if (next != blocks_.end() && !next->second.used && next->second.slab == it->second.slab && next->first == it->first + it->second.size) {
        drop_free(next->first, next->second.size);
        it->second.size += next->second.size;
        blocks_.erase(next);
    }
    if (it != blocks_.begin()) {   // and with the preceding one
        auto prev = std::prev(it);
        if (!prev->second.used && prev->second.slab == it->second.slab && prev->first + prev->second.size == it->first) {
            drop_free(prev->first, prev->second.size);
            prev->second.size += it->second.size;
            blocks_.erase(it);
            it = prev;
        }
    }
    free_.emplace(it->second.size, it->first);
}

void CUDADeviceAllocator::release_cached_memory() const {
    std::lock_guard<std::mutex> lock(mu_);
    for (size_t s = 0; s < slabs_.size(); ++s) {
        if (!slabs_[s].first) continue;
        auto it = blocks_.find(reinterpret_cast<uintptr_t>(slabs_[s].first));
        if (it != blocks_.end() && !it->second.used && it->second.size == slabs_[s].second) {   // whole slab free
            drop_free(it->first, it->second.size);
            blocks_.erase(it);
            cudaFree(slabs_[s].first);
            slabs_[s] = {nullptr, 0};
        }
    }
}

size_t CUDADeviceAllocator::bytes_reserved() const {
    std::lock_guard<std::mutex> lock(mu_);
    size_t n = 0;
    for (auto& s : slabs_) n += s.second;
    return n;
}
size_t CUDADeviceAllocator::bytes_in_use() const {
    std::lock_guard<std::mutex> lock(mu_);
    return in_use_;
}

std::shared_ptr<CPUDeviceAllocator> CPUDeviceAllocatorFactory::get_instance() {
    static std::shared_ptr<CPUDeviceAllocator> inst = std::make_shared<CPUDeviceAllocator>();   // thread-safe init
    return inst;
}
std::shared_ptr<CUDADeviceAllocator> CUDADeviceAllocatorFactory::get_instance() {
    static std::shared_ptr<CUDADeviceAllocator> inst = std::make_shared<CUDADeviceAllocator>();
    return inst;
}

// --------------------------------------------------------------------------------------------- Buffer ----
Buffer::Buffer(size_t byte_size, std::shared_ptr<DeviceAllocator> allocator, void* ptr, bool use_external)
    : byte_size_(byte_size), ptr_(ptr), use_external_(use_external), allocator_(std::move(allocator)) {
    if (!ptr_ && allocator_) {
        device_type_ = allocator_->device_type();
        use_external_ = false;
        ptr_ = allocator_->allocate(byte_size);
    }
}

Buffer::~Buffer() {
    if (!use_external_ && ptr_ && allocator_) allocator_->release(ptr_);
}

bool Buffer::allocate() {
    if (!allocator_ || byte_size_ == 0) return false;
    use_external_ = false;
    ptr_ = allocator_->allocate(byte_size_);
    return ptr_ != nullptr;
}

static base::MemcpyKind kind_of(base::DeviceType src, base::DeviceType dst) {
    const bool s = src == base::DeviceType::kDeviceCUDA, d = dst == base::DeviceType::kDeviceCUDA;
    return s ? (d ? base::MemcpyKind::kMemcpyCUDA2CUDA : base::MemcpyKind::kMemcpyCUDA2CPU)
             : (d ? base::MemcpyKind::kMemcpyCPU2CUDA : base::MemcpyKind::kMemcpyCPU2CPU);
}

void Buffer::copy_from(const Buffer& buffer) const {
    if (!allocator_) LOG(" ERROR! alloccator is empty while coping! ");
    if (!buffer.ptr_) LOG(" ERROR! ptr is empty while coping! ");
    if (buffer.device_type() == base::DeviceType::kDeviceUnknown || device_type() == base::DeviceType::kDeviceUnknown)
        LOG(" ERROR! DeviceType is unknown! ");
    allocator_->memcpy(buffer.ptr_, ptr_, std::min(byte_size_, buffer.byte_size_), kind_of(buffer.device_type(), device_type()));
}
void Buffer::copy_from(const Buffer* buffer) const {
    if (!buffer) LOG(" ERROR! ptr is empty while coping! ");
    copy_from(*buffer);
}

// --------------------------------------------------------------------------------------------- Tensor ----
static size_t count_of(const std::vector<int32_t>& dims) {
    if (dims.empty()) return 0;
    size_t n = 1;
    for (int32_t d : dims) n *= static_cast<size_t>(d);
    return n;
}

Tensor::Tensor(std::vector<int32_t> dims, bool need_alloc, std::shared_ptr<DeviceAllocator> alloc, void* ptr, base::DataType dtype)
    : size_(count_of(dims)), dims_(std::move(dims)), dtype_(dtype) {
    if (need_alloc && alloc) allocate(alloc);
    else init_buffer(alloc, need_alloc, ptr);
}

bool Tensor::allocate(std::shared_ptr<DeviceAllocator> allocator, bool need_realloc) {
    if (!allocator) {
        std::cout << " Allocator is Empty! " << std::endl;
        return false;
    }
    if (buffer_ && byte_size() <= buffer_->byte_size() && !need_realloc) return true;
    buffer_ = std::make_shared<Buffer>(byte_size(), allocator, nullptr);
    if (!buffer_->ptr()) LOG("The memory allocated is a null pointer!");
    return true;
}

void Tensor::init_buffer(std::shared_ptr<DeviceAllocator> alloc, bool need_alloc, void* ptr) {
    if (!alloc && !need_alloc) buffer_ = std::make_shared<Buffer>(byte_size(), nullptr, ptr, true);   // external view
    else allocate(alloc, true);
}

base::DeviceType Tensor::device_type() const { return buffer_ ? buffer_->device_type() : base::DeviceType::kDeviceUnknown; }
void Tensor::set_device_type(base::DeviceType device_type) const { if (buffer_) buffer_->set_device_type(device_type); }
bool Tensor::is_empty() const { return size_ == 0 || !buffer_ || !buffer_->ptr(); }

void Tensor::to_cpu() {
    if (!buffer_) LOG(" No buffer in Tensor! ");
    const auto dev = device_type();
    if (dev == base::DeviceType::kDeviceUnknown) LOG(" The device type of the tensor is unknown. ");
    if (dev == base::DeviceType::kDeviceCPU) { std::cout << " The device type of the tensor is already cpu. " << std::endl; return; }
    auto cpu_alloc = CPUDeviceAllocatorFactory::get_instance();
    auto host = std::make_shared<Buffer>(byte_size(), cpu_alloc);
    cpu_alloc->memcpy(buffer_->ptr(), host->ptr(), byte_size(), base::MemcpyKind::kMemcpyCUDA2CPU);
    buffer_ = host;
}

void Tensor::to_cuda() {
    if (!buffer_) LOG(" No buffer in Tensor! ");
    const auto dev = device_type();
    if (dev == base::DeviceType::kDeviceUnknown) LOG(" The device type of the tensor is unknown. ");
    if (dev == base::DeviceType::kDeviceCUDA) { std::cout << " The device type of the tensor is already cuda. " << std::endl; return; }
    auto cu_alloc = CUDADeviceAllocatorFactory::get_instance();
    auto devbuf = std::make_shared<Buffer>(byte_size(), cu_alloc);
    cu_alloc->memcpy(buffer_->ptr(), devbuf->ptr(), byte_size(), base::MemcpyKind::kMemcpyCPU2CUDA);
    buffer_ = devbuf;
}

void Tensor::reshape(const std::vector<int32_t>& dims) {
    const size_t n = count_of(dims);
    if (buffer_ && n > size_) {
        auto grown = std::make_shared<Buffer>(n * base::data_type_size(dtype_), buffer_->allocator());
        grown->copy_from(buffer_.get());
        buffer_ = grown;
    }
    dims_ = dims;
    size_ = n;
}

int32_t Tensor::get_dim(int32_t idx) const {
    if (idx < 0 || idx >= dims_size()) LOG("idx is wrong!");
    return dims_.at(idx);
}

bool Tensor::assign(std::shared_ptr<Buffer> buffer) {
    if (!buffer) { std::cout << "The buffer parameter in the assign function is null pointer!" << std::endl; return false; }
    if (buffer_ && buffer_->device_type() != buffer->device_type()) {
        std::cout << "The device type of the new buffer is different from the original one." << std::endl;
        return false;
    }
    if (byte_size() > buffer->byte_size()) { std::cout << "The size of buffer is too small for the tensor!" << std::endl; return false; }
    buffer_ = std::move(buffer);
    return true;
}

void Tensor::reset(const std::vector<int32_t>& dims) {
    dims_ = dims;
    size_ = count_of(dims);
    buffer_ = nullptr;
}

std::vector<size_t> Tensor::strides() const {
    std::vector<size_t> out(dims_.size(), 1);
    for (int i = static_cast<int>(dims_.size()) - 2; i >= 0; --i) out[i] = out[i + 1] * static_cast<size_t>(dims_[i + 1]);
    return out;
}

Tensor Tensor::clone() const {
    Tensor t = *this;
    t.buffer_ = std::make_shared<Buffer>(byte_size(), buffer_->allocator());
    t.buffer_->copy_from(buffer_.get());
    return t;
}

std::pair<Tensor, Tensor> slice_KV_cache(int32_t layer_idx, int32_t pos, int32_t max_seq_len, int32_t dim, const Tensor& key_cache,
                                         const Tensor& value_cache) {
    const int64_t off = (static_cast<int64_t>(layer_idx) * max_seq_len + pos) * dim;
    Tensor key({dim}, false, nullptr, const_cast<float*>(key_cache.ptr<float>(off)));
    Tensor value({dim}, false, nullptr, const_cast<float*>(value_cache.ptr<float>(off)));
    key.set_device_type(key_cache.device_type());
    value.set_device_type(value_cache.device_type());
    return {key, value};
}

}  // namespace mem
