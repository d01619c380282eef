#include "fdpass.h"

#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>

#include "../core/log.h"

namespace istore::fabric {
namespace {

constexpr int kMaxFds = 16;

// abstract namespace: no file on disk, vanishes with the process
socklen_t make_addr(const std::string& name, sockaddr_un* sa) {
    std::memset(sa, 0, sizeof(*sa));
    sa->sun_family = AF_UNIX;
    const size_t n = std::min(name.size(), sizeof(sa->sun_path) - 2);
    std::memcpy(sa->sun_path + 1, name.data(), n);
    return socklen_t(offsetof(sockaddr_un, sun_path) + 1 + n);
}

bool send_fds(int sock, const std::vector<int>& fds, uint64_t payload) {
    uint64_t hdr[2] = {fds.size(), payload};
    iovec iov{hdr, sizeof(hdr)};
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int) * kMaxFds)];
    std::memset(ctrl, 0, sizeof(ctrl));
    msghdr mh{};
    mh.msg_iov = &iov;
    mh.msg_iovlen = 1;
    if (!fds.empty()) {
        mh.msg_control = ctrl;
        mh.msg_controllen = CMSG_SPACE(sizeof(int) * fds.size());
        cmsghdr* c = CMSG_FIRSTHDR(&mh);
        c->cmsg_level = SOL_SOCKET;
        c->cmsg_type = SCM_RIGHTS;
        c->cmsg_len = CMSG_LEN(sizeof(int) * fds.size());
        std::memcpy(CMSG_DATA(c), fds.data(), sizeof(int) * fds.size());
    }
    return sendmsg(sock, &mh, MSG_NOSIGNAL) == ssize_t(sizeof(hdr));
}

bool recv_fds(int sock, std::vector<int>* fds, uint64_t* payload) {
    uint64_t hdr[2] = {0, 0};
    iovec iov{hdr, sizeof(hdr)};
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int) * kMaxFds)];
    msghdr mh{};
    mh.msg_iov = &iov;
    mh.msg_iovlen = 1;
    mh.msg_control = ctrl;
    mh.msg_controllen = sizeof(ctrl);
    if (recvmsg(sock, &mh, MSG_CMSG_CLOEXEC) != ssize_t(sizeof(hdr))) return false;
    fds->clear();
    for (cmsghdr* c = CMSG_FIRSTHDR(&mh); c; c = CMSG_NXTHDR(&mh, c)) {
        if (c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) continue;
        const size_t n = (c->cmsg_len - CMSG_LEN(0)) / sizeof(int);
        const int* p = reinterpret_cast<const int*>(CMSG_DATA(c));
        fds->insert(fds->end(), p, p + n);
    }
    *payload = hdr[1];
    return fds->size() == hdr[0];
}

}  // namespace

bool FdServer::start(const std::string& name, Provider provider, std::string* err) {
    name_ = name;
    provider_ = std::move(provider);
    listen_fd_ = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    sockaddr_un sa;
    const socklen_t len = make_addr(name, &sa);
    if (listen_fd_ < 0 || bind(listen_fd_, reinterpret_cast<sockaddr*>(&sa), len) != 0 ||
        listen(listen_fd_, 64) != 0) {
        if (err) *err = std::string("fd server '") + name + "': " + std::strerror(errno);
        if (listen_fd_ >= 0) close(listen_fd_);
        listen_fd_ = -1;
        return false;
    }
    stop_.store(false);
    thread_ = std::thread([this] { loop(); });
    return true;
}

void FdServer::stop() {
    stop_.store(true);
    if (thread_.joinable()) thread_.join();
    if (listen_fd_ >= 0) close(listen_fd_);
    listen_fd_ = -1;
}

void FdServer::loop() {
    while (!stop_.load()) {
        pollfd p{listen_fd_, POLLIN, 0};
        if (poll(&p, 1, 200) <= 0) continue;
        const int c = accept4(listen_fd_, nullptr, nullptr, SOCK_CLOEXEC);
        if (c < 0) continue;
        int32_t request = -1;
        timeval tv{2, 0};
        setsockopt(c, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
        if (recv(c, &request, sizeof(request), MSG_WAITALL) == ssize_t(sizeof(request))) {
            std::vector<int> fds;
            uint64_t payload = 0;
            if (provider_(request, &fds, &payload))
                send_fds(c, fds, payload);
            else
                send_fds(c, {}, 0);
            for (int fd : fds) close(fd);  // the kernel duplicated them into the message
        }
        close(c);
    }
}

bool fd_request(const std::string& name, int request, std::vector<int>* fds, uint64_t* payload,
                std::string* err) {
    const int s = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    sockaddr_un sa;
    const socklen_t len = make_addr(name, &sa);
    timeval tv{5, 0};
    if (s >= 0) setsockopt(s, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    const int32_t req = request;
    bool ok = s >= 0 && connect(s, reinterpret_cast<sockaddr*>(&sa), len) == 0 &&
              send(s, &req, sizeof(req), MSG_NOSIGNAL) == ssize_t(sizeof(req)) &&
              recv_fds(s, fds, payload);
    if (!ok && err) *err = std::string("fd request to '") + name + "': " + std::strerror(errno);
    if (s >= 0) close(s);
    return ok;
}

}  // namespace istore::fabric
