// Passing file descriptors between the store server and its clients (SCM_RIGHTS over an
// abstract unix-domain socket).  CUDA VMM allocations and multicast objects - unlike legacy
// cudaMalloc memory - are shared across processes as POSIX fds, not as 64-byte IPC handles,
// so the NVLS replica segment needs this side channel beside the TCP control plane.
#pragma once

#include <atomic>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace istore::fabric {

// Serves `provider(request)` -> fds to every connecting client.  The request is one int32
// (the client's CUDA device ordinal); the reply is the fds plus a u64 payload.
class FdServer {
   public:
    using Provider = std::function<bool(int request, std::vector<int>* fds, uint64_t* payload)>;
    FdServer() = default;
    ~FdServer() { stop(); }
    bool start(const std::string& name, Provider provider, std::string* err);
    void stop();
    const std::string& name() const { return name_; }

   private:
    void loop();
    std::string name_;
    Provider provider_;
    int listen_fd_ = -1;
    std::thread thread_;
    std::atomic<bool> stop_{false};
};

// Client side: connect to `name`, send `request`, receive the fds (caller closes them).
bool fd_request(const std::string& name, int request, std::vector<int>* fds, uint64_t* payload,
                std::string* err);

}  // namespace istore::fabric
