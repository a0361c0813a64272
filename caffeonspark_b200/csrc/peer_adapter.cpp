// peer_adapter.cpp -- see peer_adapter.hpp.
#include "peer_adapter.hpp"

#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <stdio.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/un.h>
#include <stdlib.h>
#include <unistd.h>

#include <chrono>
#include <random>

namespace cosb {
namespace {

constexpr uint32_t kMagic = 0x42534f43u;  // "COSB"
enum MsgType : uint32_t { kHello = 1, kCtrl = 2, kFetch = 3, kFetchReply = 4 };

// 16-byte frame header; the reference's is {rank, type, size} = 12 bytes
// (socket.cpp:30-36).
struct Header {
  uint32_t magic;
  uint32_t type;
  int32_t rank;
  uint32_t len;
};

bool send_all(int fd, const void* buf, size_t len, int pass_fd) {
  const char* p = static_cast<const char*>(buf);
  bool first = true;
  while (len > 0) {
    struct msghdr msg;
    memset(&msg, 0, sizeof(msg));
    struct iovec iov;
    iov.iov_base = const_cast<char*>(p);
    iov.iov_len = len;
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    char cbuf[CMSG_SPACE(sizeof(int))];
    if (first && pass_fd >= 0) {
      memset(cbuf, 0, sizeof(cbuf));
      msg.msg_control = cbuf;
      msg.msg_controllen = sizeof(cbuf);
      struct cmsghdr* c = CMSG_FIRSTHDR(&msg);
      c->cmsg_level = SOL_SOCKET;
      c->cmsg_type = SCM_RIGHTS;
      c->cmsg_len = CMSG_LEN(sizeof(int));
      memcpy(CMSG_DATA(c), &pass_fd, sizeof(int));
    }
    ssize_t n = sendmsg(fd, &msg, MSG_NOSIGNAL);
    if (n < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    first = false;
    p += n;
    len -= static_cast<size_t>(n);
  }
  return true;
}

// Reads exactly len bytes; any descriptor that arrives as ancillary data is
// returned in *got_fd (else left untouched).  Returns false on EOF/error.
bool recv_all(int fd, void* buf, size_t len, int* got_fd) {
  char* p = static_cast<char*>(buf);
  while (len > 0) {
    struct msghdr msg;
    memset(&msg, 0, sizeof(msg));
    struct iovec iov;
    iov.iov_base = p;
    iov.iov_len = len;
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    char cbuf[CMSG_SPACE(sizeof(int))];
    msg.msg_control = cbuf;
    msg.msg_controllen = sizeof(cbuf);
    ssize_t n = recvmsg(fd, &msg, MSG_CMSG_CLOEXEC);
    if (n < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    if (n == 0) return false;  // EOF (the reference spins here: socket.cpp:57-63)
    for (struct cmsghdr* c = CMSG_FIRSTHDR(&msg); c; c = CMSG_NXTHDR(&msg, c)) {
      if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
        int f;
        memcpy(&f, CMSG_DATA(c), sizeof(int));
        if (got_fd) *got_fd = f; else close(f);
      }
    }
    p += n;
    len -= static_cast<size_t>(n);
  }
  return true;
}

bool send_msg(int fd, uint32_t type, int rank, const std::string& payload, int pass_fd) {
  Header h{kMagic, type, rank, static_cast<uint32_t>(payload.size())};
  std::string frame(reinterpret_cast<const char*>(&h), sizeof(h));
  frame += payload;
  return send_all(fd, frame.data(), frame.size(), pass_fd);
}

bool recv_msg(int fd, Header* h, std::string* payload, int* got_fd) {
  if (!recv_all(fd, h, sizeof(*h), got_fd)) return false;
  if (h->magic != kMagic || h->len > (1u << 20)) return false;
  payload->resize(h->len);
  if (h->len && !recv_all(fd, &(*payload)[0], h->len, got_fd)) return false;
  return true;
}

void set_timeouts(int fd, int ms) {
  struct timeval tv;
  tv.tv_sec = ms / 1000;
  tv.tv_usec = (ms % 1000) * 1000;
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
}

// Endpoint names: "cosb200-..." lives in the abstract namespace (no file, vanishes with the process; needs a
// shared NETWORK namespace); a name starting with '/' is a pathname socket (COS_SOCKET_DIR, for executors in
// separate containers that share a directory).
socklen_t endpoint_addr(const std::string& name, struct sockaddr_un* sa) {
  memset(sa, 0, sizeof(*sa));
  sa->sun_family = AF_UNIX;
  size_t n = name.size();
  if (!name.empty() && name[0] == '/') {
    if (n > sizeof(sa->sun_path) - 1) n = sizeof(sa->sun_path) - 1;
    memcpy(sa->sun_path, name.data(), n);
    return static_cast<socklen_t>(offsetof(struct sockaddr_un, sun_path) + n + 1);
  }
  if (n > sizeof(sa->sun_path) - 2) n = sizeof(sa->sun_path) - 2;
  memcpy(sa->sun_path + 1, name.data(), n);  // sun_path[0] == 0 -> abstract namespace
  return static_cast<socklen_t>(offsetof(struct sockaddr_un, sun_path) + 1 + n);
}

}  // namespace

PeerAdapter::PeerAdapter(int cluster_size, int rank)
    : cluster_size_(cluster_size), rank_(rank) {
  out_fd_.assign(cluster_size_, -1);
  out_mu_.resize(cluster_size_, nullptr);
  for (int i = 0; i < cluster_size_; ++i) out_mu_[i] = new std::mutex();
  peer_pid_.assign(cluster_size_, -1);
  ctrl_recv_.assign(cluster_size_, 0);

  listen_fd_ = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (listen_fd_ < 0) {
    init_error_ = std::string("socket(AF_UNIX): ") + strerror(errno);
    return;
  }
  std::random_device rd;
  char leaf[96];
  snprintf(leaf, sizeof(leaf), "cosb200-%ld-r%d-%08x%08x", static_cast<long>(getpid()), rank_, rd(), rd());
  std::string name = leaf;
  if (const char* dir = getenv("COS_SOCKET_DIR")) {
    if (dir[0] == '/') {
      name = std::string(dir) + "/" + leaf + ".sock";
      path_ = name;
      unlink(path_.c_str());
    }
  }
  struct sockaddr_un sa;
  if (name.size() > sizeof(sa.sun_path) - 2) {
    init_error_ = "socket path too long: " + name;
    close(listen_fd_);
    listen_fd_ = -1;
    return;
  }
  socklen_t slen = endpoint_addr(name, &sa);
  if (bind(listen_fd_, reinterpret_cast<struct sockaddr*>(&sa), slen) < 0 || listen(listen_fd_, 64) < 0) {
    init_error_ = std::string("bind/listen: ") + strerror(errno);
    close(listen_fd_);
    listen_fd_ = -1;
    return;
  }
  address_ = "cosb200://" + std::to_string(static_cast<long>(getpid())) + "/" + name;
  listener_ = std::thread(&PeerAdapter::listen_loop, this);
}

PeerAdapter::~PeerAdapter() {
  close_all();
  if (listener_.joinable()) listener_.join();
  std::vector<std::thread> servers;
  {
    std::lock_guard<std::mutex> g(mu_);
    servers.swap(servers_);
  }
  for (auto& t : servers)
    if (t.joinable()) t.join();
  for (auto& kv : offers_)
    if (kv.second.fd >= 0) close(kv.second.fd);
  for (auto* m : out_mu_) delete m;
  if (!path_.empty()) unlink(path_.c_str());
}

void PeerAdapter::close_all() {
  std::lock_guard<std::mutex> g(mu_);
  stop_ = true;
  if (listen_fd_ >= 0) shutdown(listen_fd_, SHUT_RDWR);
  for (int fd : server_fds_)
    if (fd >= 0) shutdown(fd, SHUT_RDWR);
  for (int& fd : out_fd_) {
    if (fd >= 0) {
      shutdown(fd, SHUT_RDWR);
      close(fd);
      fd = -1;
    }
  }
  cv_.notify_all();
}

bool PeerAdapter::parse_address(const std::string& addr, long* pid, std::string* name) {
  static const char kScheme[] = "cosb200://";
  if (addr.compare(0, sizeof(kScheme) - 1, kScheme) != 0) return false;
  size_t slash = addr.find('/', sizeof(kScheme) - 1);
  if (slash == std::string::npos || slash + 1 >= addr.size()) return false;
  std::string pid_s = addr.substr(sizeof(kScheme) - 1, slash - (sizeof(kScheme) - 1));
  if (pid_s.empty() || pid_s.find_first_not_of("0123456789") != std::string::npos) return false;
  *pid = atol(pid_s.c_str());
  *name = addr.substr(slash + 1);
  if (!name->empty() && (*name)[0] == '/') {  // pathname socket: .../cosb200-<pid>-r<rank>-<nonce>.sock
    size_t leaf = name->find_last_of('/');
    return name->compare(leaf + 1, 8, "cosb200-") == 0;
  }
  return name->compare(0, 8, "cosb200-") == 0;
}

void PeerAdapter::listen_loop() {
  for (;;) {
    int fd = accept4(listen_fd_, nullptr, nullptr, SOCK_CLOEXEC);
    if (fd < 0) {
      if (errno == EINTR) continue;
      break;  // shut down
    }
    std::lock_guard<std::mutex> g(mu_);
    if (stop_) {
      close(fd);
      break;
    }
    server_fds_.push_back(fd);
    servers_.emplace_back(&PeerAdapter::serve, this, fd);
  }
  close(listen_fd_);
}

// One server thread per incoming channel (socket.cpp:79-127 has one receiver
// pthread per peer).  Requests are answered on the same connection.
void PeerAdapter::serve(int fd) {
  int src = -1;
  // The rendezvous name is discoverable (/proc/net/unix) and a fetch hands out a read/write handle of the GPU
  // arena: only serve peers running under our own effective uid (or root).  COS_ALLOW_ANY_UID=1 lifts this for
  // executors whose containers map uids differently.
  {
    struct ucred cred;
    socklen_t len = sizeof(cred);
    const bool known = getsockopt(fd, SOL_SOCKET, SO_PEERCRED, &cred, &len) == 0 && len == sizeof(cred);
    if (!getenv("COS_ALLOW_ANY_UID") && (!known || (cred.uid != geteuid() && cred.uid != 0))) {
      shutdown(fd, SHUT_RDWR);
      close(fd);
      std::lock_guard<std::mutex> g(mu_);
      for (int& f : server_fds_)
        if (f == fd) f = -1;
      return;
    }
  }
  for (;;) {
    Header h;
    std::string payload;
    if (!recv_msg(fd, &h, &payload, nullptr)) break;
    if (h.type == kHello) {
      if (h.rank < 0 || h.rank >= cluster_size_) break;
      src = h.rank;
      char buf[32];
      snprintf(buf, sizeof(buf), "%ld", static_cast<long>(getpid()));
      if (!send_msg(fd, kHello, rank_, buf, -1)) break;
    } else if (h.type == kCtrl) {
      if (src < 0) break;
      std::lock_guard<std::mutex> g(mu_);
      ctrl_recv_[src]++;
      cv_.notify_all();
    } else if (h.type == kFetch) {
      // payload = "<timeout_ms>\n<key>"
      size_t nl = payload.find('\n');
      if (nl == std::string::npos) break;
      int timeout_ms = atoi(payload.substr(0, nl).c_str());
      std::string key = payload.substr(nl + 1);
      Offer o{-1, ""};
      bool found = false;
      {
        std::unique_lock<std::mutex> g(mu_);
        found = cv_.wait_for(g, std::chrono::milliseconds(timeout_ms),
                             [&] { return stop_ || offers_.count(key) > 0; }) &&
                offers_.count(key) > 0;
        if (found) {
          o = offers_[key];
          if (o.fd >= 0) o.fd = dup(o.fd);  // our own copy, taken under the lock: offer() may close the original
        }
      }
      std::string reply(1, found ? (o.fd >= 0 ? 'F' : 'M') : 'N');
      reply += o.meta;
      const bool sent = send_msg(fd, kFetchReply, rank_, reply, found ? o.fd : -1);
      if (o.fd >= 0) close(o.fd);
      if (!sent) break;
    } else {
      break;
    }
  }
  shutdown(fd, SHUT_RDWR);
  close(fd);
  std::lock_guard<std::mutex> g(mu_);
  for (int& f : server_fds_)
    if (f == fd) f = -1;
}

bool PeerAdapter::connect(const std::vector<std::string>& addrs, std::string* err) {
  if (!ok()) {
    *err = "adapter not listening: " + init_error_;
    return false;
  }
  if (static_cast<int>(addrs.size()) < cluster_size_) {
    *err = "connect: expected " + std::to_string(cluster_size_) + " addresses, got " + std::to_string(addrs.size());
    return false;
  }
  // Start at rank+1 "to avoid all sending to same peer at the same time"
  // (socket_sync_cpu.cpp:137-138).
  for (int n = 1; n < cluster_size_; ++n) {
    int peer = (rank_ + n) % cluster_size_;
    long pid = 0;
    std::string name;
    if (!parse_address(addrs[peer], &pid, &name)) {
      *err = "connect: malformed address for rank " + std::to_string(peer) + ": '" + addrs[peer] + "'";
      return false;
    }
    int fd = -1;
    int backoff_ms = 20;
    for (int attempt = 0; attempt < 6 && fd < 0; ++attempt) {
      int s = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
      if (s < 0) break;
      struct sockaddr_un sa;
      socklen_t slen = endpoint_addr(name, &sa);
      if (::connect(s, reinterpret_cast<struct sockaddr*>(&sa), slen) == 0) {
        fd = s;
      } else {
        close(s);
        usleep(backoff_ms * 1000);
        backoff_ms *= 2;
      }
    }
    if (fd < 0) {
      *err = "connect: cannot reach rank " + std::to_string(peer) + " at '" + addrs[peer] + "': " + strerror(errno);
      return false;
    }
    set_timeouts(fd, 60000);
    Header h;
    std::string payload;
    if (!send_msg(fd, kHello, rank_, "", -1) || !recv_msg(fd, &h, &payload, nullptr) || h.type != kHello ||
        h.rank != peer) {
      close(fd);
      *err = "connect: handshake with rank " + std::to_string(peer) + " failed";
      return false;
    }
    std::lock_guard<std::mutex> g(mu_);
    out_fd_[peer] = fd;
    peer_pid_[peer] = atol(payload.c_str());
  }
  connected_ = true;
  return true;
}

bool PeerAdapter::barrier(int timeout_ms, std::string* err) {
  if (cluster_size_ == 1) return true;
  if (!connected_) {
    *err = "barrier: not connected";
    return false;
  }
  uint64_t want;
  {
    std::lock_guard<std::mutex> g(mu_);
    want = ++ctrl_sent_;
  }
  for (int n = 1; n < cluster_size_; ++n) {
    int peer = (rank_ + n) % cluster_size_;
    std::lock_guard<std::mutex> g(*out_mu_[peer]);
    if (out_fd_[peer] < 0 || !send_msg(out_fd_[peer], kCtrl, rank_, "", -1)) {
      *err = "barrier: send to rank " + std::to_string(peer) + " failed";
      return false;
    }
  }
  std::unique_lock<std::mutex> g(mu_);
  bool done = cv_.wait_for(g, std::chrono::milliseconds(timeout_ms), [&] {
    if (stop_) return true;
    for (int p = 0; p < cluster_size_; ++p)
      if (p != rank_ && ctrl_recv_[p] < want) return false;
    return true;
  });
  if (!done || stop_) {
    *err = "barrier: timed out after " + std::to_string(timeout_ms) + " ms waiting for peers";
    return false;
  }
  return true;
}

void PeerAdapter::offer(const std::string& key, int fd, const std::string& meta) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = offers_.find(key);
  if (it != offers_.end() && it->second.fd >= 0) close(it->second.fd);
  offers_[key] = Offer{fd >= 0 ? fcntl(fd, F_DUPFD_CLOEXEC, 0) : -1, meta};
  cv_.notify_all();
}

bool PeerAdapter::fetch(int peer, const std::string& key, int* fd, std::string* meta, int timeout_ms,
                        std::string* err) {
  if (fd) *fd = -1;
  if (peer < 0 || peer >= cluster_size_ || peer == rank_ || out_fd_[peer] < 0) {
    *err = "fetch: no channel to rank " + std::to_string(peer);
    return false;
  }
  std::lock_guard<std::mutex> g(*out_mu_[peer]);
  set_timeouts(out_fd_[peer], timeout_ms + 5000);
  std::string req = std::to_string(timeout_ms) + "\n" + key;
  Header h;
  std::string payload;
  int got = -1;
  if (!send_msg(out_fd_[peer], kFetch, rank_, req, -1) || !recv_msg(out_fd_[peer], &h, &payload, &got) ||
      h.type != kFetchReply || payload.empty()) {
    if (got >= 0) close(got);
    *err = "fetch('" + key + "') from rank " + std::to_string(peer) + ": channel error";
    return false;
  }
  if (payload[0] == 'N') {
    if (got >= 0) close(got);
    *err = "fetch('" + key + "') from rank " + std::to_string(peer) + ": not offered within " +
           std::to_string(timeout_ms) + " ms";
    return false;
  }
  if (payload[0] == 'F' && got < 0) {
    *err = "fetch('" + key + "') from rank " + std::to_string(peer) + ": descriptor did not arrive";
    return false;
  }
  if (meta) *meta = payload.substr(1);
  if (fd) *fd = got; else if (got >= 0) close(got);
  return true;
}

}  // namespace cosb
