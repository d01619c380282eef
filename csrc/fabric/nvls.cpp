#include "nvls.h"

#include <cuda.h>
#include <cuda_runtime_api.h>

#include <cstring>

#include "../core/log.h"

namespace istore::fabric {
namespace {

// Driver entry points, resolved lazily (no link-time dependency on libcuda).
struct Driver {
    bool ok = false;
    std::string why;
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*,
                          unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr,
                                  unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle,
                       unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*,
                                            CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) =
        nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t,
                                 CUmemGenericAllocationHandle, size_t, size_t,
                                 unsigned long long) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*,
                                        CUmulticastGranularity_flags) = nullptr;
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle,
                                           CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*,
                                             CUmemAllocationHandleType) = nullptr;
};

template <typename F>
bool resolve(const char* name, F& fn, std::string& why) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p ||
        q != cudaDriverEntryPointSuccess) {
        (void)cudaGetLastError();
        why = std::string("driver entry point not available: ") + name;
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

const Driver& driver() {
    static Driver d = [] {
        Driver r;
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
            (void)cudaGetLastError();
            r.why = "no CUDA device";
            return r;
        }
        cudaFree(nullptr);  // make sure the primary context exists
        bool ok = resolve("cuDeviceGet", r.DeviceGet, r.why) &&
                  resolve("cuDeviceGetAttribute", r.DeviceGetAttribute, r.why) &&
                  resolve("cuGetErrorString", r.GetErrorString, r.why) &&
                  resolve("cuMemCreate", r.MemCreate, r.why) &&
                  resolve("cuMemRelease", r.MemRelease, r.why) &&
                  resolve("cuMemAddressReserve", r.MemAddressReserve, r.why) &&
                  resolve("cuMemAddressFree", r.MemAddressFree, r.why) &&
                  resolve("cuMemMap", r.MemMap, r.why) && resolve("cuMemUnmap", r.MemUnmap, r.why) &&
                  resolve("cuMemSetAccess", r.MemSetAccess, r.why) &&
                  resolve("cuMemGetAllocationGranularity", r.MemGetAllocationGranularity, r.why) &&
                  resolve("cuMemExportToShareableHandle", r.MemExportToShareableHandle, r.why) &&
                  resolve("cuMemImportFromShareableHandle", r.MemImportFromShareableHandle, r.why);
        if (ok) {
            // multicast entry points are optional (absent before driver 535)
            std::string w;
            resolve("cuMulticastCreate", r.MulticastCreate, w);
            resolve("cuMulticastAddDevice", r.MulticastAddDevice, w);
            resolve("cuMulticastBindMem", r.MulticastBindMem, w);
            resolve("cuMulticastUnbind", r.MulticastUnbind, w);
            resolve("cuMulticastGetGranularity", r.MulticastGetGranularity, w);
        }
        r.ok = ok;
        return r;
    }();
    return d;
}

std::string cu_err(const Driver& d, CUresult r) {
    const char* s = nullptr;
    if (d.GetErrorString && d.GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "CUresult " + std::to_string(int(r));
}

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

NvlsProbe nvls_probe(int device) {
    NvlsProbe p;
    const Driver& d = driver();
    p.driver_ok = d.ok;
    if (!d.ok) {
        p.detail = d.why;
        return p;
    }
    CUdevice dev;
    if (d.DeviceGet(&dev, device) != CUDA_SUCCESS) {
        p.detail = "cuDeviceGet failed";
        return p;
    }
    int v = 0;
    if (d.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev) ==
        CUDA_SUCCESS)
        p.vmm_supported = v != 0;
    if (d.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED,
                             dev) == CUDA_SUCCESS)
        p.posix_fd_supported = v != 0;
    if (d.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev) ==
        CUDA_SUCCESS)
        p.fabric_handle_supported = v != 0;
    if (d.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS)
        p.multicast_supported = v != 0 && d.MulticastCreate != nullptr;
    if (p.multicast_supported && d.MulticastGetGranularity) {
        CUmulticastObjectProp prop{};
        prop.numDevices = 2;
        prop.size = 2u << 20;
        size_t g = 0;
        if (d.MulticastGetGranularity(&g, &prop, CU_MULTICAST_GRANULARITY_RECOMMENDED) ==
            CUDA_SUCCESS)
            p.granularity = g;
    }
    p.detail = "ok";
    return p;
}

std::shared_ptr<NvlsGroup> NvlsGroup::create(const std::vector<int>& devices, size_t bytes,
                                             std::string* err) {
    auto fail = [&](const std::string& m) -> std::shared_ptr<NvlsGroup> {
        if (err) *err = m;
        return nullptr;
    };
    const Driver& d = driver();
    if (!d.ok) return fail(d.why);
    if (!d.MulticastCreate || !d.MulticastAddDevice || !d.MulticastBindMem)
        return fail("driver has no multicast API");
    if (devices.size() < 2) return fail("a multicast group needs at least two devices");
    for (int dv : devices) {
        NvlsProbe p = nvls_probe(dv);
        if (!p.multicast_supported)
            return fail("device " + std::to_string(dv) + " does not support multicast");
    }
    std::shared_ptr<NvlsGroup> g(new NvlsGroup());
    g->devs_ = devices;

    CUmulticastObjectProp mprop{};
    mprop.numDevices = unsigned(devices.size());
    mprop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;  // shareable with clients
    size_t gran = 0;
    mprop.size = bytes;
    if (d.MulticastGetGranularity(&gran, &mprop, CU_MULTICAST_GRANULARITY_RECOMMENDED) !=
            CUDA_SUCCESS ||
        gran == 0)
        gran = 512u << 20;
    const size_t size = round_up(bytes, gran);
    mprop.size = size;
    g->bytes_ = size;

    int prev = -1;
    cudaGetDevice(&prev);
    CUresult r;
    CUmemGenericAllocationHandle mc = 0;
    if ((r = d.MulticastCreate(&mc, &mprop)) != CUDA_SUCCESS)
        return fail("cuMulticastCreate: " + cu_err(d, r));
    g->mc_handle_ = mc;
    for (int dv : devices) {
        CUdevice cd;
        d.DeviceGet(&cd, dv);
        cudaSetDevice(dv);
        cudaFree(nullptr);
        if ((r = d.MulticastAddDevice(mc, cd)) != CUDA_SUCCESS) {
            if (prev >= 0) cudaSetDevice(prev);
            return fail("cuMulticastAddDevice: " + cu_err(d, r));
        }
    }
    std::vector<CUmemAccessDesc> access(devices.size());
    for (size_t i = 0; i < devices.size(); ++i) {
        access[i].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        access[i].location.id = devices[i];
        access[i].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    }
    for (size_t i = 0; i < devices.size(); ++i) {
        CUmemAllocationProp prop{};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = devices[i];
        prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        CUmemGenericAllocationHandle mem = 0;
        if ((r = d.MemCreate(&mem, size, &prop, 0)) != CUDA_SUCCESS) {
            if (prev >= 0) cudaSetDevice(prev);
            return fail("cuMemCreate: " + cu_err(d, r));
        }
        g->mem_handles_.push_back(mem);
        if ((r = d.MulticastBindMem(mc, 0, mem, 0, size, 0)) != CUDA_SUCCESS) {
            if (prev >= 0) cudaSetDevice(prev);
            return fail("cuMulticastBindMem: " + cu_err(d, r));
        }
        CUdeviceptr va = 0;
        if ((r = d.MemAddressReserve(&va, size, gran, 0, 0)) != CUDA_SUCCESS ||
            (r = d.MemMap(va, size, 0, mem, 0)) != CUDA_SUCCESS ||
            (r = d.MemSetAccess(va, size, access.data(), access.size())) != CUDA_SUCCESS) {
            if (prev >= 0) cudaSetDevice(prev);
            return fail("mapping the local replica: " + cu_err(d, r));
        }
        g->uc_va_.push_back(uint64_t(va));
    }
    CUdeviceptr mva = 0;
    if ((r = d.MemAddressReserve(&mva, size, gran, 0, 0)) != CUDA_SUCCESS ||
        (r = d.MemMap(mva, size, 0, mc, 0)) != CUDA_SUCCESS ||
        (r = d.MemSetAccess(mva, size, access.data(), access.size())) != CUDA_SUCCESS) {
        if (prev >= 0) cudaSetDevice(prev);
        return fail("mapping the multicast object: " + cu_err(d, r));
    }
    g->mc_va_ = uint64_t(mva);
    g->mc_mapped_ = true;
    if (prev >= 0) cudaSetDevice(prev);
    LOG_INFO("NVLS group over %zu GPUs, %zu MiB replica each", devices.size(), size >> 20);
    return g;
}

int NvlsGroup::index_of_device(int device) const {
    for (size_t i = 0; i < devs_.size(); ++i)
        if (devs_[i] == device) return int(i);
    return -1;
}

static int export_fd(uint64_t handle) {
    const Driver& d = driver();
    if (!d.ok || !handle) return -1;
    int fd = -1;
    if (d.MemExportToShareableHandle(&fd, CUmemGenericAllocationHandle(handle),
                                     CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS)
        return -1;
    return fd;
}

int NvlsGroup::export_mc_fd() const { return export_fd(mc_handle_); }

int NvlsGroup::export_mem_fd(size_t i) const {
    return i < mem_handles_.size() ? export_fd(mem_handles_[i]) : -1;
}

std::shared_ptr<NvlsImport> NvlsImport::import(int mc_fd, int mem_fd, size_t bytes, int device,
                                               std::string* err) {
    auto fail = [&](const std::string& m) -> std::shared_ptr<NvlsImport> {
        if (err) *err = m;
        return nullptr;
    };
    const Driver& d = driver();
    if (!d.ok) return fail(d.why);
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    cudaFree(nullptr);
    std::shared_ptr<NvlsImport> im(new NvlsImport());
    im->bytes_ = bytes;
    CUmemAccessDesc access{};
    access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    access.location.id = device;
    access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CUresult r;
    auto map_fd = [&](int fd, uint64_t* handle_out, uint64_t* va_out) -> bool {
        CUmemGenericAllocationHandle h = 0;
        if ((r = d.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(uintptr_t(fd)),
                                                CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR)) !=
            CUDA_SUCCESS)
            return false;
        *handle_out = h;
        CUdeviceptr va = 0;
        if ((r = d.MemAddressReserve(&va, bytes, 2u << 20, 0, 0)) != CUDA_SUCCESS ||
            (r = d.MemMap(va, bytes, 0, h, 0)) != CUDA_SUCCESS ||
            (r = d.MemSetAccess(va, bytes, &access, 1)) != CUDA_SUCCESS)
            return false;
        *va_out = uint64_t(va);
        return true;
    };
    bool ok = map_fd(mc_fd, &im->mc_handle_, &im->mc_va_);
    if (ok && mem_fd >= 0) ok = map_fd(mem_fd, &im->mem_handle_, &im->uc_va_);
    if (prev >= 0) cudaSetDevice(prev);
    if (!ok) return fail("importing the multicast group: " + cu_err(d, r));
    return im;
}

NvlsImport::~NvlsImport() {
    const Driver& d = driver();
    if (!d.ok) return;
    for (uint64_t va : {mc_va_, uc_va_}) {
        if (!va) continue;
        d.MemUnmap(CUdeviceptr(va), bytes_);
        d.MemAddressFree(CUdeviceptr(va), bytes_);
    }
    if (mem_handle_) d.MemRelease(CUmemGenericAllocationHandle(mem_handle_));
    if (mc_handle_) d.MemRelease(CUmemGenericAllocationHandle(mc_handle_));
}

NvlsGroup::~NvlsGroup() {
    const Driver& d = driver();
    if (!d.ok) return;
    if (mc_mapped_) {
        d.MemUnmap(CUdeviceptr(mc_va_), bytes_);
        d.MemAddressFree(CUdeviceptr(mc_va_), bytes_);
    }
    for (size_t i = 0; i < uc_va_.size(); ++i) {
        d.MemUnmap(CUdeviceptr(uc_va_[i]), bytes_);
        d.MemAddressFree(CUdeviceptr(uc_va_[i]), bytes_);
    }
    for (uint64_t h : mem_handles_) d.MemRelease(CUmemGenericAllocationHandle(h));
    if (mc_handle_) d.MemRelease(CUmemGenericAllocationHandle(mc_handle_));
}

}  // namespace istore::fabric
