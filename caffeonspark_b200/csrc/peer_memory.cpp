// peer_memory.cpp -- see peer_memory.hpp.
#include "peer_memory.hpp"

#include <cuda.h>
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <mutex>
#include <sstream>

namespace cosb {
namespace {

// Driver entry points resolved at run time (no libcuda.so link dependency).
struct DriverApi {
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*,
                                          CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*,
                        unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemGetAccess)(unsigned long long*, const CUmemLocation*, CUdeviceptr) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                         unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*,
                                           CUmemAllocationHandleType) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t,
                               size_t, unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*,
                                      CUmulticastGranularity_flags) = nullptr;
  bool ok = false;
  std::string error;
};

template <typename F>
bool resolve(const char* name, F* fn, std::string* err) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    *err = std::string("driver entry point ") + name + " unavailable: " +
           (e != cudaSuccess ? cudaGetErrorString(e) : "not found");
    cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}

const DriverApi& driver() {
  static DriverApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    std::string e;
    bool ok = resolve("cuGetErrorString", &api.GetErrorString, &e) && resolve("cuDeviceGet", &api.DeviceGet, &e) &&
              resolve("cuDeviceGetAttribute", &api.DeviceGetAttribute, &e) &&
              resolve("cuMemGetAllocationGranularity", &api.MemGetAllocationGranularity, &e) &&
              resolve("cuMemCreate", &api.MemCreate, &e) && resolve("cuMemRelease", &api.MemRelease, &e) &&
              resolve("cuMemAddressReserve", &api.MemAddressReserve, &e) &&
              resolve("cuMemAddressFree", &api.MemAddressFree, &e) && resolve("cuMemMap", &api.MemMap, &e) &&
              resolve("cuMemUnmap", &api.MemUnmap, &e) && resolve("cuMemSetAccess", &api.MemSetAccess, &e) &&
              resolve("cuMemExportToShareableHandle", &api.MemExportToShareableHandle, &e) &&
              resolve("cuMemImportFromShareableHandle", &api.MemImportFromShareableHandle, &e);
    api.ok = ok;
    api.error = e;
    if (ok) {
      // optional (NVLS); absence just disables multicast
      std::string ignore;
      resolve("cuMemGetAccess", &api.MemGetAccess, &ignore);
      resolve("cuMulticastCreate", &api.MulticastCreate, &ignore);
      resolve("cuMulticastAddDevice", &api.MulticastAddDevice, &ignore);
      resolve("cuMulticastBindMem", &api.MulticastBindMem, &ignore);
      resolve("cuMulticastUnbind", &api.MulticastUnbind, &ignore);
      resolve("cuMulticastGetGranularity", &api.MulticastGetGranularity, &ignore);
    }
  });
  return api;
}

std::string cu_err(const char* what, CUresult r) {
  const char* s = nullptr;
  if (driver().GetErrorString) driver().GetErrorString(r, &s);
  std::ostringstream os;
  os << what << " failed: " << (s ? s : "unknown") << " (" << static_cast<int>(r) << ")";
  return os.str();
}

std::string rt_err(const char* what, cudaError_t e) {
  std::ostringstream os;
  os << what << " failed: " << cudaGetErrorString(e);
  cudaGetLastError();
  return os.str();
}

size_t round_up(size_t v, size_t g) { return (v + g - 1) / g * g; }

CUmemAllocationProp device_prop(int device) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

// Several in-process ranks on the same GPU may ask for access to the same arena at the same time:
// serialise, and skip the call when the device can already read and write the range.
std::mutex g_access_mu;

bool set_rw(CUdeviceptr va, size_t bytes, int device, std::string* err) {
  std::lock_guard<std::mutex> g(g_access_mu);
  if (driver().MemGetAccess) {
    CUmemLocation loc;
    memset(&loc, 0, sizeof(loc));
    loc.type = CU_MEM_LOCATION_TYPE_DEVICE;
    loc.id = device;
    unsigned long long flags = 0;
    if (driver().MemGetAccess(&flags, &loc, va) == CUDA_SUCCESS && flags == CU_MEM_ACCESS_FLAGS_PROT_READWRITE)
      return true;
  }
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CUresult r = driver().MemSetAccess(va, bytes, &acc, 1);
  if (r != CUDA_SUCCESS) {
    *err = cu_err("cuMemSetAccess", r);
    return false;
  }
  return true;
}

int device_attr(int device, CUdevice_attribute a) {
  const DriverApi& d = driver();
  if (!d.ok) return 0;
  CUdevice dev;
  int v = 0;
  if (d.DeviceGet(&dev, device) != CUDA_SUCCESS) return 0;
  if (d.DeviceGetAttribute(&v, a, dev) != CUDA_SUCCESS) return 0;
  return v;
}

}  // namespace

// ------------------------------------------------------------- DeviceArena

DeviceArena::~DeviceArena() { destroy(); }

bool DeviceArena::create(int device, size_t bytes, bool prefer_vmm, std::string* err) {
  destroy();
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    *err = rt_err("cudaSetDevice", e);
    return false;
  }
  e = cudaFree(0);  // make sure the primary context exists and is current
  if (e != cudaSuccess) {
    *err = rt_err("cudaFree(0)", e);
    return false;
  }
  device_ = device;
  std::string vmm_err;
  const DriverApi& d = driver();
  if (prefer_vmm && d.ok && device_attr(device, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED) &&
      device_attr(device, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED)) {
    CUmemAllocationProp prop = device_prop(device);
    size_t gran = 0;
    CUresult r = d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (r == CUDA_SUCCESS && gran > 0) {
      size_t sz = round_up(bytes, gran);
      CUmemGenericAllocationHandle h = 0;
      CUdeviceptr va = 0;
      int fd = -1;
      bool mapped = false;
      do {
        r = d.MemCreate(&h, sz, &prop, 0);
        if (r != CUDA_SUCCESS) { vmm_err = cu_err("cuMemCreate", r); h = 0; break; }
        r = d.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
        if (r != CUDA_SUCCESS) { vmm_err = cu_err("cuMemExportToShareableHandle", r); fd = -1; break; }
        r = d.MemAddressReserve(&va, sz, gran, 0, 0);
        if (r != CUDA_SUCCESS) { vmm_err = cu_err("cuMemAddressReserve", r); va = 0; break; }
        r = d.MemMap(va, sz, 0, h, 0);
        if (r != CUDA_SUCCESS) { vmm_err = cu_err("cuMemMap", r); break; }
        mapped = true;
        if (!set_rw(va, sz, device, &vmm_err)) break;
        vmm_ = true;
      } while (0);
      if (vmm_) {
        base_ = reinterpret_cast<void*>(va);
        bytes_ = sz;
        handle_ = h;
        fd_ = fd;
        transport_ = kTransportVmmFd;
      } else {
        if (mapped) d.MemUnmap(va, sz);
        if (va) d.MemAddressFree(va, sz);
        if (fd >= 0) ::close(fd);
        if (h) d.MemRelease(h);
      }
    } else {
      vmm_err = cu_err("cuMemGetAllocationGranularity", r);
    }
  } else if (prefer_vmm) {
    vmm_err = d.ok ? "device lacks VMM / POSIX-fd export support" : d.error;
  }
  if (!vmm_) {
    // legacy path: cudaMalloc + cudaIpcGetMemHandle
    size_t sz = round_up(bytes, 2u << 20);
    void* p = nullptr;
    e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) {
      *err = rt_err("cudaMalloc", e) + (vmm_err.empty() ? "" : " (VMM path: " + vmm_err + ")");
      return false;
    }
    base_ = p;
    bytes_ = sz;
    transport_ = kTransportLegacyIpc;
  }
  e = cudaMemset(base_, 0, bytes_);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    *err = rt_err("cudaMemset(arena)", e);
    destroy();
    return false;
  }
  return true;
}

void DeviceArena::destroy() {
  if (!base_) return;
  cudaSetDevice(device_);
  if (vmm_) {
    const DriverApi& d = driver();
    CUdeviceptr va = reinterpret_cast<CUdeviceptr>(base_);
    d.MemUnmap(va, bytes_);
    d.MemAddressFree(va, bytes_);
    d.MemRelease(handle_);
    if (fd_ >= 0) ::close(fd_);
  } else {
    cudaFree(base_);
  }
  base_ = nullptr;
  bytes_ = 0;
  fd_ = -1;
  handle_ = 0;
  vmm_ = false;
}

uint64_t process_nonce() {
  static const uint64_t nonce = [] {
    uint64_t v = 0;
    if (FILE* f = fopen("/dev/urandom", "rb")) {
      if (fread(&v, sizeof(v), 1, f) != 1) v = 0;
      fclose(f);
    }
    if (v == 0) {
      struct timespec ts;
      clock_gettime(CLOCK_REALTIME, &ts);
      v = (static_cast<uint64_t>(getpid()) << 40) ^ (static_cast<uint64_t>(ts.tv_sec) << 20) ^
          static_cast<uint64_t>(ts.tv_nsec) ^ reinterpret_cast<uint64_t>(&v);
    }
    return v | 1;  // never 0
  }();
  return nonce;
}

ArenaMeta DeviceArena::meta() const {
  ArenaMeta m;
  memset(&m, 0, sizeof(m));
  m.version = 2;
  m.proc_nonce = process_nonce();
  m.transport = transport_;
  m.pid = static_cast<int64_t>(getpid());
  m.device = device_;
  m.bytes = bytes_;
  m.base_ptr = reinterpret_cast<uint64_t>(base_);
  if (!vmm_ && base_) {
    cudaIpcMemHandle_t h;
    cudaSetDevice(device_);
    if (cudaIpcGetMemHandle(&h, base_) == cudaSuccess) {
      static_assert(sizeof(h) == sizeof(m.ipc_handle), "ipc handle size");
      memcpy(m.ipc_handle, &h, sizeof(h));
    } else {
      cudaGetLastError();
    }
  }
  return m;
}

bool DeviceArena::grant_access(int other_device, std::string* err) {
  if (other_device == device_) return true;
  if (vmm_) return set_rw(reinterpret_cast<CUdeviceptr>(base_), bytes_, other_device, err);
  // cudaMalloc memory: classic peer access, enabled from the accessing device
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(other_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(device_, 0);
  cudaSetDevice(prev);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
    *err = rt_err("cudaDeviceEnablePeerAccess", e);
    return false;
  }
  cudaGetLastError();
  return true;
}

// ------------------------------------------------------------- PeerMapping

PeerMapping::~PeerMapping() { close(); }

bool PeerMapping::open(const ArenaMeta& m, int fd, int my_device, std::string* err) {
  close();
  if (m.version != 2) {
    if (fd >= 0) ::close(fd);
    *err = "peer arena metadata version mismatch";
    return false;
  }
  cudaError_t e = cudaSetDevice(my_device);
  if (e != cudaSuccess) {
    if (fd >= 0) ::close(fd);
    *err = rt_err("cudaSetDevice", e);
    return false;
  }
  if (m.proc_nonce == process_nonce()) {  // NOT the raw pid: pids collide across PID namespaces
    // in-process rank: same address space (the owner grants access if needed)
    if (fd >= 0) ::close(fd);
    if (m.device != my_device) {  // another GPU of this process: open the access path from here
      if (m.transport == kTransportVmmFd) {
        if (!set_rw(static_cast<CUdeviceptr>(m.base_ptr), m.bytes, my_device, err)) return false;
      } else {
        e = cudaDeviceEnablePeerAccess(m.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          *err = rt_err("cudaDeviceEnablePeerAccess", e);
          return false;
        }
        cudaGetLastError();
      }
    }
    base_ = reinterpret_cast<void*>(m.base_ptr);
    bytes_ = m.bytes;
    kind_ = 0;
    return true;
  }
  if (m.transport == kTransportVmmFd) {
    const DriverApi& d = driver();
    if (!d.ok) {
      if (fd >= 0) ::close(fd);
      *err = d.error;
      return false;
    }
    if (fd < 0) {
      *err = "peer arena uses VMM transport but no descriptor was received";
      return false;
    }
    CUmemGenericAllocationHandle h = 0;
    CUresult r = d.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                                CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    ::close(fd);
    if (r != CUDA_SUCCESS) {
      *err = cu_err("cuMemImportFromShareableHandle", r);
      return false;
    }
    CUdeviceptr va = 0;
    r = d.MemAddressReserve(&va, m.bytes, 2u << 20, 0, 0);
    if (r != CUDA_SUCCESS) {
      d.MemRelease(h);
      *err = cu_err("cuMemAddressReserve(peer)", r);
      return false;
    }
    r = d.MemMap(va, m.bytes, 0, h, 0);
    if (r != CUDA_SUCCESS) {
      d.MemAddressFree(va, m.bytes);
      d.MemRelease(h);
      *err = cu_err("cuMemMap(peer)", r);
      return false;
    }
    if (!set_rw(va, m.bytes, my_device, err)) {
      d.MemUnmap(va, m.bytes);
      d.MemAddressFree(va, m.bytes);
      d.MemRelease(h);
      return false;
    }
    base_ = reinterpret_cast<void*>(va);
    bytes_ = m.bytes;
    handle_ = h;
    kind_ = 1;
    return true;
  }
  if (m.transport == kTransportLegacyIpc) {
    if (fd >= 0) ::close(fd);
    cudaIpcMemHandle_t h;
    memcpy(&h, m.ipc_handle, sizeof(h));
    void* p = nullptr;
    e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      *err = rt_err("cudaIpcOpenMemHandle", e);
      return false;
    }
    base_ = p;
    bytes_ = m.bytes;
    kind_ = 2;
    return true;
  }
  if (fd >= 0) ::close(fd);
  *err = "unknown peer arena transport";
  return false;
}

void PeerMapping::close() {
  if (!base_) return;
  if (kind_ == 1) {
    const DriverApi& d = driver();
    CUdeviceptr va = reinterpret_cast<CUdeviceptr>(base_);
    d.MemUnmap(va, bytes_);
    d.MemAddressFree(va, bytes_);
    d.MemRelease(handle_);
  } else if (kind_ == 2) {
    cudaIpcCloseMemHandle(base_);
    cudaGetLastError();
  }
  base_ = nullptr;
  bytes_ = 0;
  kind_ = -1;
  handle_ = 0;
}

// -------------------------------------------------------- MulticastMapping

MulticastMapping::~MulticastMapping() { close(); }

bool MulticastMapping::supported(int device) {
  const DriverApi& d = driver();
  return d.ok && d.MulticastCreate && d.MulticastAddDevice && d.MulticastBindMem &&
         device_attr(device, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED) != 0;
}

bool MulticastMapping::create(int device, size_t bytes, int ndevices, int* fd_out, std::string* err) {
  const DriverApi& d = driver();
  if (!supported(device)) {
    *err = "multicast not supported on this device/driver";
    return false;
  }
  cudaSetDevice(device);
  CUmulticastObjectProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.numDevices = static_cast<unsigned>(ndevices);
  prop.size = bytes;
  prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  if (d.MulticastGetGranularity &&
      d.MulticastGetGranularity(&gran, &prop, CU_MULTICAST_GRANULARITY_MINIMUM) == CUDA_SUCCESS && gran) {
    if (bytes % gran) {
      *err = "arena size " + std::to_string(bytes) + " is not a multiple of the multicast granularity " +
             std::to_string(gran);
      return false;
    }
  }
  CUmemGenericAllocationHandle h = 0;
  CUresult r = d.MulticastCreate(&h, &prop);
  if (r != CUDA_SUCCESS) {
    *err = cu_err("cuMulticastCreate", r);
    return false;
  }
  int fd = -1;
  r = d.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    d.MemRelease(h);
    *err = cu_err("cuMemExportToShareableHandle(multicast)", r);
    return false;
  }
  handle_ = h;
  bytes_ = bytes;
  device_ = device;
  *fd_out = fd;
  return true;
}

bool MulticastMapping::import(int device, size_t bytes, int fd, std::string* err) {
  const DriverApi& d = driver();
  if (!supported(device)) {
    if (fd >= 0) ::close(fd);
    *err = "multicast not supported on this device/driver";
    return false;
  }
  cudaSetDevice(device);
  CUmemGenericAllocationHandle h = 0;
  CUresult r = d.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                              CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  ::close(fd);
  if (r != CUDA_SUCCESS) {
    *err = cu_err("cuMemImportFromShareableHandle(multicast)", r);
    return false;
  }
  handle_ = h;
  bytes_ = bytes;
  device_ = device;
  return true;
}

bool MulticastMapping::add_device(std::string* err) {
  const DriverApi& d = driver();
  CUdevice dev;
  CUresult r = d.DeviceGet(&dev, device_);
  if (r == CUDA_SUCCESS) r = d.MulticastAddDevice(handle_, dev);
  if (r != CUDA_SUCCESS) {
    *err = cu_err("cuMulticastAddDevice", r);
    return false;
  }
  return true;
}

bool MulticastMapping::bind_and_map(const DeviceArena& arena, std::string* err) {
  const DriverApi& d = driver();
  if (arena.transport() != kTransportVmmFd || arena.bytes() != bytes_) {
    *err = "multicast needs a VMM arena of the multicast object's size";
    return false;
  }
  cudaSetDevice(device_);
  CUresult r = d.MulticastBindMem(handle_, 0, arena.vmm_handle(), 0, bytes_, 0);
  if (r != CUDA_SUCCESS) {
    *err = cu_err("cuMulticastBindMem", r);
    return false;
  }
  bound_ = true;
  CUdeviceptr va = 0;
  r = d.MemAddressReserve(&va, bytes_, 2u << 20, 0, 0);
  if (r != CUDA_SUCCESS) {
    *err = cu_err("cuMemAddressReserve(multicast)", r);
    return false;
  }
  r = d.MemMap(va, bytes_, 0, handle_, 0);
  if (r != CUDA_SUCCESS) {
    d.MemAddressFree(va, bytes_);
    *err = cu_err("cuMemMap(multicast)", r);
    return false;
  }
  if (!set_rw(va, bytes_, device_, err)) {
    d.MemUnmap(va, bytes_);
    d.MemAddressFree(va, bytes_);
    return false;
  }
  base_ = reinterpret_cast<void*>(va);
  return true;
}

void MulticastMapping::close() {
  const DriverApi& d = driver();
  if (base_) {
    CUdeviceptr va = reinterpret_cast<CUdeviceptr>(base_);
    d.MemUnmap(va, bytes_);
    d.MemAddressFree(va, bytes_);
    base_ = nullptr;
  }
  if (handle_) {
    if (bound_ && d.MulticastUnbind) {
      CUdevice dev;
      if (d.DeviceGet(&dev, device_) == CUDA_SUCCESS) d.MulticastUnbind(handle_, dev, 0, bytes_);
    }
    d.MemRelease(handle_);
    handle_ = 0;
  }
  bound_ = false;
}

std::string peer_memory_capabilities(int device) {
  std::ostringstream os;
  const DriverApi& d = driver();
  os << "driver_api=" << (d.ok ? "ok" : ("unavailable(" + d.error + ")"));
  if (d.ok) {
    os << " vmm=" << device_attr(device, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED)
       << " posix_fd=" << device_attr(device, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED)
       << " multicast=" << device_attr(device, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED);
  }
  return os.str();
}

}  // namespace cosb
