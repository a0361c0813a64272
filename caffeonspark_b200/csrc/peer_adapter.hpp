// peer_adapter.hpp -- control-plane transport between executors.
//
// Plays the role of the reference's SocketAdapter + SocketChannel
// (caffe-distri/include/util/socket.hpp:22-89, src/main/cpp/util/socket.cpp):
// one listener thread per process whose address string is what
// CaffeNet.localAddresses() returns and what Spark broadcasts
// (caffe-grid .../CaffeOnSpark.scala:113-154), one outgoing channel per peer.
// Unlike the reference, channels never carry tensor payload: they move
//   * CTRL tokens   -- the zero-payload barrier of SocketSync::sync(false)
//                      (socket_sync_cpu.cpp:135-163),
//   * memory handles -- POSIX file descriptors of CUDA VMM allocations
//                      (SCM_RIGHTS) plus a small metadata blob,
// after which all gradient/weight traffic is NVLink peer memory.
#ifndef COS_PEER_ADAPTER_HPP_
#define COS_PEER_ADAPTER_HPP_

#include <condition_variable>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace cosb {

class PeerAdapter {
 public:
  PeerAdapter(int cluster_size, int rank);
  ~PeerAdapter();
  PeerAdapter(const PeerAdapter&) = delete;
  PeerAdapter& operator=(const PeerAdapter&) = delete;

  bool ok() const { return listen_fd_ >= 0; }
  const std::string& init_error() const { return init_error_; }
  int cluster_size() const { return cluster_size_; }
  int rank() const { return rank_; }

  // "cosb200://<pid>/<endpoint>" (reference: "host:port", socket.hpp:28-37).  <endpoint> is an abstract
  // unix-socket name, or an absolute socket path when COS_SOCKET_DIR is set.
  const std::string& address() const { return address_; }
  static bool parse_address(const std::string& addr, long* pid, std::string* name);

  // Connect one outgoing channel per peer (socket.cpp:242-281 retries with
  // back-off; ours retries 6 times from 20 ms).  addrs is indexed by rank,
  // addrs[rank()] is ignored.  False on a malformed/unreachable address.
  bool connect(const std::vector<std::string>& addrs, std::string* err);
  bool connected() const { return connected_; }
  long peer_pid(int peer) const { return peer_pid_[peer]; }

  // CTRL barrier across all ranks.
  bool barrier(int timeout_ms, std::string* err);

  // Publish (key -> fd, meta) for peers to fetch.  fd may be -1 (meta only).
  // The adapter dup()s the fd; the caller keeps ownership of its copy.
  void offer(const std::string& key, int fd, const std::string& meta);
  // Fetch what `peer` offered under `key`; blocks (peer side) until offered or
  // timeout.  *fd receives a new descriptor (or -1).
  bool fetch(int peer, const std::string& key, int* fd, std::string* meta, int timeout_ms, std::string* err);

 private:
  struct Offer {
    int fd;
    std::string meta;
  };
  void listen_loop();
  void serve(int fd);
  void close_all();

  const int cluster_size_;
  const int rank_;
  std::string address_, init_error_;
  std::string path_;  // pathname socket to unlink (COS_SOCKET_DIR), empty for the abstract namespace
  int listen_fd_ = -1;
  bool stop_ = false;
  bool connected_ = false;

  std::thread listener_;
  std::mutex mu_;  // guards everything below
  std::condition_variable cv_;
  std::vector<std::thread> servers_;
  std::vector<int> server_fds_;
  std::vector<int> out_fd_;               // outgoing channel per peer
  std::vector<std::mutex*> out_mu_;       // per-channel send mutex (socket.hpp:66)
  std::vector<long> peer_pid_;
  std::vector<uint64_t> ctrl_recv_;       // CTRL tokens received per source rank
  uint64_t ctrl_sent_ = 0;
  std::map<std::string, Offer> offers_;
};

}  // namespace cosb
#endif
