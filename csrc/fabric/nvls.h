// NVLS multicast groups: one-writer / many-reader replication through the NVSwitch.
//
// Not present in the reference (it serves a shared prefix with N unicast RDMA reads,
// src/infinistore.cpp:424-533).  A group owns, per GPU, a VMM allocation (the local
// replica) bound to one multicast object; a store to the multicast VA (multimem.st in
// kernels/kv_bcast_nvls.cu) lands in every replica.  The driver API is resolved at run
// time through cudaGetDriverEntryPoint, so the module still imports on CPU-only hosts.
#pragma once

#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace istore::fabric {

struct NvlsProbe {
    bool driver_ok = false;
    bool multicast_supported = false;
    bool vmm_supported = false;
    bool posix_fd_supported = false;
    bool fabric_handle_supported = false;
    size_t granularity = 0;
    std::string detail;
};
NvlsProbe nvls_probe(int device);

class NvlsGroup {
   public:
    ~NvlsGroup();
    // All devices are driven by this process (the store server owning several pool GPUs,
    // or a single-process test).  `bytes` is rounded up to the multicast granularity.
    static std::shared_ptr<NvlsGroup> create(const std::vector<int>& devices, size_t bytes,
                                             std::string* err);
    // POSIX file descriptors of the multicast object / of replica i, to be passed to another
    // process over a unix socket (fabric/fdpass.h).  The caller closes them.  -1 on failure.
    int export_mc_fd() const;
    int export_mem_fd(size_t i) const;
    const std::vector<int>& devices() const { return devs_; }
    int index_of_device(int device) const;
    uint64_t mc_ptr(size_t i) const { return i < devs_.size() ? mc_va_ : 0; }
    uint64_t uc_ptr(size_t i) const { return i < uc_va_.size() ? uc_va_[i] : 0; }
    size_t bytes() const { return bytes_; }
    size_t size() const { return devs_.size(); }

   private:
    NvlsGroup() = default;
    std::vector<int> devs_;
    size_t bytes_ = 0;
    uint64_t mc_handle_ = 0;
    uint64_t mc_va_ = 0;
    std::vector<uint64_t> mem_handles_;
    std::vector<uint64_t> uc_va_;
    bool mc_mapped_ = false;
};

// A client process' view of a group owned by another process: the multicast object mapped
// for writing and (optionally) the replica of the client's own GPU mapped for reading.
class NvlsImport {
   public:
    ~NvlsImport();
    static std::shared_ptr<NvlsImport> import(int mc_fd, int mem_fd, size_t bytes, int device,
                                              std::string* err);
    uint64_t mc_ptr() const { return mc_va_; }
    uint64_t uc_ptr() const { return uc_va_; }
    size_t bytes() const { return bytes_; }

   private:
    NvlsImport() = default;
    size_t bytes_ = 0;
    uint64_t mc_handle_ = 0, mem_handle_ = 0;
    uint64_t mc_va_ = 0, uc_va_ = 0;
};

}  // namespace istore::fabric
