#include "segment.h"

#include <cuda_runtime_api.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <random>

#include "../core/log.h"

namespace istore::fabric {
namespace {

std::mutex g_err_mu;
std::string g_last_error;

void set_error(const std::string& e) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_last_error = e;
}

bool cuda_ok(cudaError_t e, const char* what, std::string* err) {
    if (e == cudaSuccess) return true;
    std::string msg = std::string(what) + ": " + cudaGetErrorString(e);
    (void)cudaGetLastError();  // clear the sticky-less error state
    set_error(msg);
    if (err) *err = msg;
    return false;
}

struct Uuid {
    uint8_t b[16];
    Uuid() {
        std::random_device rd;
        for (int i = 0; i < 16; i += 4) {
            const uint32_t r = rd();
            std::memcpy(b + i, &r, 4);
        }
        const uint32_t pid = uint32_t(getpid());
        std::memcpy(b, &pid, 4);  // make collisions between live processes impossible
    }
};

// Set the device for the duration of a scope and restore the caller's device afterwards
// (the module must not disturb torch's notion of the current device).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (dev >= 0 && dev != prev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

constexpr size_t kAlign = 2u << 20;  // 2 MiB: TLB page granularity of the GPU MMU
size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

const uint8_t* process_uuid() {
    static Uuid u;
    return u.b;
}

int cuda_device_count() {
    static int n = [] {
        int c = 0;
        if (cudaGetDeviceCount(&c) != cudaSuccess) {
            (void)cudaGetLastError();
            c = 0;
        }
        return c;
    }();
    return n;
}

bool cuda_available() { return cuda_device_count() > 0; }

std::string cuda_last_error() {
    std::lock_guard<std::mutex> lk(g_err_mu);
    return g_last_error;
}

// ---------------------------------------------------------------- SegmentOwner

std::unique_ptr<SegmentOwner> SegmentOwner::create_host(uint32_t id, size_t bytes,
                                                        uint32_t granule, int port,
                                                        std::string* err) {
    std::unique_ptr<SegmentOwner> s(new SegmentOwner());
    char name[64];
    std::snprintf(name, sizeof(name), "/istore_b200_%d_%d_%u", int(getpid()), port, id);
    shm_unlink(name);
    const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) {
        if (err) *err = std::string("shm_open failed: ") + std::strerror(errno);
        return nullptr;
    }
    const size_t map_bytes = round_up(bytes, 4096);
    if (ftruncate(fd, off_t(map_bytes)) != 0) {
        if (err) *err = std::string("ftruncate failed: ") + std::strerror(errno);
        close(fd);
        shm_unlink(name);
        return nullptr;
    }
    void* p = mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) {
        if (err) *err = std::string("mmap failed: ") + std::strerror(errno);
        close(fd);
        shm_unlink(name);
        return nullptr;
    }
    s->base_ = p;
    s->shm_fd_ = fd;
    s->shm_name_ = name;
    SegmentInfo& i = s->info_;
    i.id = id;
    i.kind = kSegHostShm;
    i.device = -1;
    i.granule = granule;
    i.bytes = bytes;
    i.index_off = 0;
    i.index_slots = 0;
    i.map_bytes = map_bytes;
    std::memset(i.handle, 0, sizeof(i.handle));
    std::memcpy(i.handle, name, std::strlen(name));
    std::memcpy(i.owner, process_uuid(), 16);
    i.owner_ptr = reinterpret_cast<uint64_t>(p);
    return s;
}

std::unique_ptr<SegmentOwner> SegmentOwner::create_device(uint32_t id, int device, size_t bytes,
                                                          uint32_t granule, size_t index_slots,
                                                          std::string* err) {
    if (device < 0 || device >= cuda_device_count()) {
        if (err) *err = "no such CUDA device: " + std::to_string(device);
        return nullptr;
    }
    DeviceGuard g(device);
    if (!g.ok) {
        if (err) *err = "cudaSetDevice failed";
        return nullptr;
    }
    std::unique_ptr<SegmentOwner> s(new SegmentOwner());
    const size_t index_off = round_up(bytes, kAlign);
    const size_t map_bytes = round_up(index_off + index_slots * 32, kAlign);
    void* p = nullptr;
    if (!cuda_ok(cudaMalloc(&p, map_bytes), "cudaMalloc(pool segment)", err)) return nullptr;
    s->base_ = p;
    SegmentInfo& i = s->info_;
    i.id = id;
    i.kind = kSegDeviceIpc;
    i.device = device;
    i.granule = granule;
    i.bytes = bytes;
    i.index_off = index_slots ? index_off : 0;
    i.index_slots = index_slots;
    i.map_bytes = map_bytes;
    std::memcpy(i.owner, process_uuid(), 16);
    i.owner_ptr = reinterpret_cast<uint64_t>(p);
    if (index_slots &&
        !cuda_ok(cudaMemset(static_cast<uint8_t*>(p) + index_off, 0, index_slots * 32),
                 "cudaMemset(index)", err))
        return nullptr;
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == sizeof(i.handle), "IPC handle is 64 bytes");
    if (!cuda_ok(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle", err)) {
        // Not fatal: same-process clients (SPMD ranks hosting their own shard) still work.
        LOG_WARN("pool segment %u is not exportable over CUDA IPC: %s", id,
                 err ? err->c_str() : "");
        std::memset(i.handle, 0, sizeof(i.handle));
        if (err) err->clear();
    } else {
        std::memcpy(i.handle, &h, sizeof(h));
    }
    cudaDeviceSynchronize();
    return s;
}

std::unique_ptr<SegmentOwner> SegmentOwner::create_replica(uint32_t id,
                                                           const std::vector<int>& devices,
                                                           size_t bytes, uint32_t granule, int port,
                                                           std::string* err) {
    std::unique_ptr<SegmentOwner> s(new SegmentOwner());
    s->group_ = NvlsGroup::create(devices, bytes, err);
    if (!s->group_) return nullptr;
    char name[64];
    std::snprintf(name, sizeof(name), "istore_b200_nvls_%d_%d_%u", int(getpid()), port, id);
    s->fd_server_ = std::make_unique<FdServer>();
    std::shared_ptr<NvlsGroup> group = s->group_;
    auto provider = [group](int device, std::vector<int>* fds, uint64_t* payload) {
        const int mc = group->export_mc_fd();
        if (mc < 0) return false;
        fds->push_back(mc);
        const int idx = group->index_of_device(device);
        if (idx >= 0) {
            const int mem = group->export_mem_fd(size_t(idx));
            if (mem >= 0) fds->push_back(mem);
        }
        *payload = group->bytes();
        return true;
    };
    if (!s->fd_server_->start(name, provider, err)) return nullptr;
    SegmentInfo& i = s->info_;
    i.id = id;
    i.kind = kSegReplica;
    i.device = kReplicaDevice;
    i.granule = granule;
    i.bytes = s->group_->bytes() / granule * granule;
    i.index_off = 0;
    i.index_slots = 0;
    i.map_bytes = s->group_->bytes();
    std::memset(i.handle, 0, sizeof(i.handle));
    std::memcpy(i.handle, name, std::strlen(name));
    std::memcpy(i.owner, process_uuid(), 16);
    i.owner_ptr = reinterpret_cast<uint64_t>(s->group_.get());
    s->base_ = reinterpret_cast<void*>(s->group_->mc_ptr(0));
    return s;
}

void SegmentOwner::clear_index() {
    if (info_.kind != kSegDeviceIpc || !info_.index_slots) return;
    DeviceGuard g(info_.device);
    std::string err;
    cuda_ok(cudaMemset(static_cast<uint8_t*>(base_) + info_.index_off, 0, info_.index_slots * 32),
            "cudaMemset(index)", &err);
    cudaDeviceSynchronize();
}

SegmentOwner::~SegmentOwner() {
    if (info_.kind == kSegReplica) {
        fd_server_.reset();
        group_.reset();
        return;
    }
    if (!base_) return;
    if (info_.kind == kSegHostShm) {
        munmap(base_, info_.map_bytes);
        if (shm_fd_ >= 0) close(shm_fd_);
        if (!shm_name_.empty()) shm_unlink(shm_name_.c_str());
    } else {
        DeviceGuard g(info_.device);
        cudaFree(base_);
    }
}

// ---------------------------------------------------------------- client mappings

Mapping::~Mapping() {
    if (info.kind == kSegHostShm) {
        if (host_registered) {
            DeviceGuard g(device);
            cudaHostUnregister(host_ptr);
        }
        if (host_ptr && std::memcmp(info.owner, process_uuid(), 16) != 0)
            munmap(host_ptr, info.map_bytes);
    } else if (ipc_opened && dev_ptr) {
        DeviceGuard g(device);
        cudaIpcCloseMemHandle(dev_ptr);
    }
}

namespace {
struct MapKey {
    uint8_t owner[16];
    uint32_t id;
    int device;
    uint64_t owner_ptr;
    bool operator<(const MapKey& o) const { return std::memcmp(this, &o, sizeof(MapKey)) < 0; }
};
std::mutex g_map_mu;
std::map<MapKey, std::weak_ptr<Mapping>> g_maps;
}  // namespace

std::shared_ptr<Mapping> map_segment(const SegmentInfo& info, int device, std::string* err) {
    MapKey key;
    std::memset(&key, 0, sizeof(key));
    std::memcpy(key.owner, info.owner, 16);
    key.id = info.id;
    key.device = device;
    key.owner_ptr = info.owner_ptr;
    const bool same_process = std::memcmp(info.owner, process_uuid(), 16) == 0;
    std::lock_guard<std::mutex> lk(g_map_mu);
    // Only real mappings are cached (an IPC handle can be opened once per context).  A
    // segment of this very process is used through its own pointer; caching that would let
    // a mapping of a stopped server alias a new segment that reuses the address.
    if (!same_process) {
        auto it = g_maps.find(key);
        if (it != g_maps.end()) {
            if (auto sp = it->second.lock()) return sp;
            g_maps.erase(it);
        }
    }
    auto m = std::make_shared<Mapping>();
    m->info = info;
    m->device = device;

    if (info.kind == kSegHostShm) {
        if (same_process) {
            m->host_ptr = reinterpret_cast<uint8_t*>(info.owner_ptr);
        } else {
            char name[65];
            std::memcpy(name, info.handle, 64);
            name[64] = 0;
            const int fd = shm_open(name, O_RDWR, 0600);
            if (fd < 0) {
                if (err) *err = std::string("shm_open(") + name + "): " + std::strerror(errno);
                return nullptr;
            }
            void* p = mmap(nullptr, info.map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            close(fd);
            if (p == MAP_FAILED) {
                if (err) *err = std::string("mmap(pool): ") + std::strerror(errno);
                return nullptr;
            }
            m->host_ptr = static_cast<uint8_t*>(p);
        }
        if (device >= 0) {
            // let kernels on `device` reach the host tier over PCIe
            DeviceGuard g(device);
            std::string e;
            if (cuda_ok(cudaHostRegister(m->host_ptr, info.map_bytes,
                                         cudaHostRegisterMapped | cudaHostRegisterPortable),
                        "cudaHostRegister(pool)", &e)) {
                void* dp = nullptr;
                if (cuda_ok(cudaHostGetDevicePointer(&dp, m->host_ptr, 0),
                            "cudaHostGetDevicePointer", &e)) {
                    m->dev_ptr = static_cast<uint8_t*>(dp);
                    m->host_registered = true;
                }
            } else if (e.find("already") != std::string::npos) {
                void* dp = nullptr;  // another mapping of this process registered it
                if (cudaHostGetDevicePointer(&dp, m->host_ptr, 0) == cudaSuccess)
                    m->dev_ptr = static_cast<uint8_t*>(dp);
                (void)cudaGetLastError();
            }
        }
    } else if (info.kind == kSegReplica) {
        if (device < 0) {
            if (err) *err = "the NVLS-replicated region needs a CUDA device on the client";
            return nullptr;
        }
        if (same_process) {
            auto* group = reinterpret_cast<NvlsGroup*>(info.owner_ptr);
            const int idx = group->index_of_device(device);
            m->mc_ptr = reinterpret_cast<uint8_t*>(group->mc_ptr(0));
            m->dev_ptr = reinterpret_cast<uint8_t*>(group->uc_ptr(idx >= 0 ? size_t(idx) : 0));
        } else {
            char name[65];
            std::memcpy(name, info.handle, 64);
            name[64] = 0;
            std::vector<int> fds;
            uint64_t bytes = 0;
            if (!fd_request(name, device, &fds, &bytes, err)) return nullptr;
            if (fds.empty()) {
                if (err) *err = "the server exported no multicast handle";
                return nullptr;
            }
            m->nvls = NvlsImport::import(fds[0], fds.size() > 1 ? fds[1] : -1, size_t(bytes), device,
                                         err);
            for (int fd : fds) close(fd);
            if (!m->nvls) return nullptr;
            m->mc_ptr = reinterpret_cast<uint8_t*>(m->nvls->mc_ptr());
            m->dev_ptr = reinterpret_cast<uint8_t*>(m->nvls->uc_ptr());  // null: no local replica
        }
    } else {
        if (device < 0) {
            if (err) *err = "an HBM pool segment needs a CUDA device on the client";
            return nullptr;
        }
        DeviceGuard g(device);
        if (same_process) {
            m->dev_ptr = reinterpret_cast<uint8_t*>(info.owner_ptr);
            if (info.device != device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, device, info.device);
                if (!can) {
                    if (err) *err = "no peer access between client and pool device";
                    return nullptr;
                }
                const cudaError_t e = cudaDeviceEnablePeerAccess(info.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    cuda_ok(e, "cudaDeviceEnablePeerAccess", err);
                    return nullptr;
                }
                (void)cudaGetLastError();
            }
        } else {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, info.handle, sizeof(h));
            void* p = nullptr;
            if (!cuda_ok(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess),
                         "cudaIpcOpenMemHandle(pool)", err))
                return nullptr;
            m->dev_ptr = static_cast<uint8_t*>(p);
            m->ipc_opened = true;
        }
    }
    if (!same_process) g_maps[key] = m;
    return m;
}

}  // namespace istore::fabric
