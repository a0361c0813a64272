// peer_memory.hpp -- device memory that other executors can map.
//
// Replaces the reference's per-peer staging buffers and wire transfers
// (SocketBuffer, socket.cpp:340-406: cudaMemcpy D2H -> TCP -> H2D; RDMABuffer,
// rdma.cpp:340-378: ibv RDMA write) with ONE peer-mappable arena per rank that
// holds data_, diff_ and the barrier flags.  Peers map it once at connect()
// time; afterwards the kernels load/store it directly over NVLink.
//
// Sharing mechanisms, in order of preference:
//   same process  : raw pointer (in-process ranks; cuMemSetAccess for other devices)
//   VMM           : cuMemCreate + POSIX fd export, fd sent over the PeerAdapter
//                   (SCM_RIGHTS), cuMemImportFromShareableHandle + cuMemMap
//   legacy IPC    : cudaMalloc + cudaIpcGetMemHandle (64-byte handle in metadata)
// The driver API is reached through cudaGetDriverEntryPoint so the library has
// no link-time dependency on libcuda.so (it must load on a GPU-less box).
#ifndef COS_PEER_MEMORY_HPP_
#define COS_PEER_MEMORY_HPP_

#include <cstddef>
#include <cstdint>
#include <string>

namespace cosb {

enum ArenaTransport : int32_t { kTransportVmmFd = 0, kTransportLegacyIpc = 1 };

// What a peer needs to map an arena; sent as the metadata blob next to the fd.
struct ArenaMeta {
  uint32_t version;
  int32_t transport;
  int64_t pid;
  int32_t device;
  int32_t reserved;
  uint64_t bytes;
  uint64_t base_ptr;  // valid inside the process identified by (pid, proc_nonce) only
  // Random per-process token (process_nonce()).  PIDs collide across PID namespaces (executors in separate
  // containers reached through COS_SOCKET_DIR often all run as PID 1), so "same process" is decided by the
  // nonce; the pid is kept for diagnostics.
  uint64_t proc_nonce;
  unsigned char ipc_handle[64];
};

// 64 random bits drawn once per process (from /dev/urandom; falls back to pid, clock and an address).
uint64_t process_nonce();

class DeviceArena {
 public:
  DeviceArena() = default;
  ~DeviceArena();
  DeviceArena(const DeviceArena&) = delete;
  DeviceArena& operator=(const DeviceArena&) = delete;

  // Allocates `bytes` (rounded up to the allocation granularity) on `device`,
  // zero-filled.  prefer_vmm=false forces the legacy cudaMalloc/IPC path.
  bool create(int device, size_t bytes, bool prefer_vmm, std::string* err);
  void destroy();

  void* base() const { return base_; }
  size_t bytes() const { return bytes_; }
  int device() const { return device_; }
  int transport() const { return transport_; }
  // Exported descriptor (VMM) or -1.  Owned by the arena.
  int fd() const { return fd_; }
  uint64_t vmm_handle() const { return handle_; }
  ArenaMeta meta() const;
  // Let another device of this process access the arena (in-process ranks).
  bool grant_access(int other_device, std::string* err);

 private:
  void* base_ = nullptr;
  size_t bytes_ = 0;
  int device_ = -1;
  int transport_ = kTransportVmmFd;
  int fd_ = -1;
  uint64_t handle_ = 0;  // CUmemGenericAllocationHandle
  bool vmm_ = false;
};

// A peer's arena mapped into this process / device.
class PeerMapping {
 public:
  PeerMapping() = default;
  ~PeerMapping();
  PeerMapping(const PeerMapping&) = delete;
  PeerMapping& operator=(const PeerMapping&) = delete;

  // `fd` is consumed (closed) when the transport needs it.
  bool open(const ArenaMeta& meta, int fd, int my_device, std::string* err);
  void close();
  void* base() const { return base_; }
  size_t bytes() const { return bytes_; }

 private:
  void* base_ = nullptr;
  size_t bytes_ = 0;
  int kind_ = -1;  // 0 = same-process pointer, 1 = VMM import, 2 = legacy IPC
  uint64_t handle_ = 0;
};

// NVLS multicast object spanning all ranks' arenas (where the platform exposes
// it): cuMulticastCreate on rank 0, handle exported as fd, every rank adds its
// device and binds its arena, then maps the multicast handle.
class MulticastMapping {
 public:
  MulticastMapping() = default;
  ~MulticastMapping();
  static bool supported(int device);
  // Rank 0: create the object for `ndevices`; returns exportable fd.
  bool create(int device, size_t bytes, int ndevices, int* fd_out, std::string* err);
  // Other ranks: import rank 0's fd (consumed).
  bool import(int device, size_t bytes, int fd, std::string* err);
  bool add_device(std::string* err);
  // Bind the local arena (must be VMM) and map the multicast VA.
  bool bind_and_map(const DeviceArena& arena, std::string* err);
  void close();
  void* base() const { return base_; }

 private:
  uint64_t handle_ = 0;
  void* base_ = nullptr;
  size_t bytes_ = 0;
  int device_ = -1;
  bool bound_ = false;
};

// Human-readable probe of what the platform offers (used by tests/diagnostics).
std::string peer_memory_capabilities(int device);

}  // namespace cosb
#endif
