// hdf5_io.hpp -- the subset of the HDF5 file format Caffe's snapshots use, written and read without libhdf5.
//
// The reference snapshots with snapshot_format: HDF5 through libhdf5's H5LT calls (net.cpp:867-917 Net::ToHDF5,
// sgd_solver.cpp:279-323 SnapshotSolverStateToHDF5, util/hdf5.cpp hdf5_save_nd_dataset / hdf5_save_int /
// hdf5_save_string) and restores with net.cpp:805-851 CopyTrainedLayersFromHDF5 and sgd_solver.cpp:325-347
// RestoreSolverStateFromHDF5.  What those calls put on disk with default property lists is the original
// ("version 0") layout of the HDF5 File Format Specification:
//   superblock v0 -> root group = v1 object header with a Symbol Table message -> group B-tree (v1, 'TREE') +
//   local heap ('HEAP') + symbol table nodes ('SNOD'); datasets = v1 object headers with Dataspace (v1),
//   Datatype (v1: IEEE float32 LE / int32 LE / fixed string), Fill Value (v2), contiguous Data Layout (v3) and
//   Modification Time messages, raw data stored contiguously.
// This file writes exactly those structures (byte layouts mirrored from libhdf5-written files: the reference's
// own fixtures caffe-public/src/caffe/test/test_data/{solver_data,sample_data}.h5) and reads them back.
// There is NO libhdf5 (nor h5py) in this build environment: the READER is pinned on those libhdf5-written fixtures
// (tests/test_hdf5_io.py), the WRITER is checked by round trips through that reader and by structural comparison
// with the fixtures; it has not been opened by libhdf5 itself -- said plainly in DESIGN.md.
// Unsupported on read (clear error, never a guess): chunked / compressed / compact layouts, superblock >= 2,
// new-style (fractal heap) groups, datatypes other than float32 / int32 / uint8-int8 / fixed strings.
#ifndef COS_HDF5_IO_HPP_
#define COS_HDF5_IO_HPP_

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace cosb {

struct H5Node {
  enum Kind { kGroup, kFloat32, kInt32, kString };
  std::string name;
  Kind kind = kGroup;
  std::vector<std::unique_ptr<H5Node>> children;  // groups, in insertion order
  std::vector<int64_t> shape;                     // datasets (strings: scalar)
  // payload: for writing, `data` may point at caller-owned memory (count elements); reading fills `f32`/`i32`/`str`
  const void* data = nullptr;
  uint64_t count = 0;
  std::vector<float> f32;
  std::vector<int32_t> i32;
  std::string str;

  H5Node* add_group(const std::string& n);
  H5Node* add_float(const std::string& n, const std::vector<int64_t>& shape, const float* data, uint64_t count);
  H5Node* add_int(const std::string& n, int32_t v);  // H5LTmake_dataset_int(loc, n, 1, {1}, &v)
  H5Node* add_string(const std::string& n, const std::string& s);  // H5LTmake_dataset_string
  const H5Node* find(const std::string& n) const;  // direct child by name
};

// Writes `root`'s children as the root group of a new file.  Groups hold at most 256 links (one B-tree level).
bool h5_write(const std::string& path, const H5Node& root, std::string* err);
// Reads the whole tree (dataset payloads included).  Children of a group come back in name order (B-tree order).
bool h5_read(const std::string& path, H5Node* root, std::string* err);
// True if the file starts with the HDF5 signature.
bool h5_is_hdf5(const std::string& path);

}  // namespace cosb
#endif
