// ref_driver.cpp -- CPU ORACLE driver (test infrastructure, NOT product code).
//
// Runs the reference's OWN inter-executor sync code -- socket.cpp,
// socket_sync_cpu.cpp and parallel_cpu.cpp compiled verbatim from
// /root/reference/caffe-distri/src/main/cpp/util/ against oracle/shim/ -- as N
// processes on localhost loopback, the way caffe_mini_cluster does
// (caffe-distri/src/main/cpp/util/mini_cluster.cpp:69-164), and wraps it with
// the restated SGD update of oracle/sync_oracle.c.  Two uses:
//   --dump DIR : write per-rank own-shard weights/history after every
//                iteration + the final all-gathered weights (golden vectors,
//                pins sync_oracle.c bit-exactly);
//   --time     : time on_start + on_gradients_ready + update per iteration
//                (the CPU baseline of bench.py, kind "reference").
// One Solver::Step (solver.cpp:194-273) is emulated as: diff := 0,
// on_start(), diff := synthetic local gradient, on_gradients_ready(),
// ApplyUpdate restatement on the FULL buffer, ++iter.
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "caffe/caffe.hpp"
#include "util/parallel_cpu.hpp"
#include "util/socket.hpp"
#include "util/socket_sync_cpu.hpp"

#include "sync_oracle.h"

namespace caffe {
// Params<Dtype>::Params lives in libcaffe's parallel.cpp:70-74 in the
// reference; restated here with the total_size rule (parallel.cpp:60-68).
template <typename Dtype>
static size_t shim_total_size(const vector<Blob<Dtype>*>& params) {
  size_t size = 0;
  for (size_t i = 0; i < params.size(); ++i) size += params[i]->count();
  return (size > 0) ? size : 1;
}
template <typename Dtype>
Params<Dtype>::Params(shared_ptr<Solver<Dtype> > root_solver)
    : size_(shim_total_size<Dtype>(root_solver->net()->learnable_params())),
      data_(),
      diff_() {}
template class Params<float>;
template class Params<double>;
}  // namespace caffe

using caffe::shared_ptr;
using std::string;
using std::vector;

struct Args {
  std::map<string, string> kv;
  string get(const string& k, const string& d) const {
    std::map<string, string>::const_iterator it = kv.find(k);
    return it == kv.end() ? d : it->second;
  }
  long geti(const string& k, long d) const {
    return kv.count(k) ? atol(kv.find(k)->second.c_str()) : d;
  }
  double getf(const string& k, double d) const {
    return kv.count(k) ? atof(kv.find(k)->second.c_str()) : d;
  }
};

template <typename T>
static vector<T> parse_list(const string& s) {
  vector<T> out;
  std::stringstream ss(s);
  string tok;
  while (std::getline(ss, tok, ',')) {
    if (tok.empty()) continue;
    out.push_back((T)atof(tok.c_str()));
  }
  return out;
}

static int policy_id(const string& p) {
  if (p == "fixed") return COS_ORACLE_LR_FIXED;
  if (p == "step") return COS_ORACLE_LR_STEP;
  if (p == "exp") return COS_ORACLE_LR_EXP;
  if (p == "inv") return COS_ORACLE_LR_INV;
  if (p == "multistep") return COS_ORACLE_LR_MULTISTEP;
  if (p == "poly") return COS_ORACLE_LR_POLY;
  if (p == "sigmoid") return COS_ORACLE_LR_SIGMOID;
  fprintf(stderr, "unknown lr_policy %s\n", p.c_str());
  exit(2);
}

static void write_file_atomic(const string& path, const string& content) {
  string tmp = path + ".tmp";
  FILE* f = fopen(tmp.c_str(), "w");
  fwrite(content.data(), 1, content.size(), f);
  fclose(f);
  rename(tmp.c_str(), path.c_str());
}

static string wait_read_file(const string& path) {
  for (int tries = 0; tries < 60000; ++tries) {
    FILE* f = fopen(path.c_str(), "r");
    if (f) {
      char buf[512];
      size_t n = fread(buf, 1, sizeof(buf) - 1, f);
      fclose(f);
      buf[n] = 0;
      return string(buf);
    }
    usleep(1000);
  }
  fprintf(stderr, "timeout waiting for %s\n", path.c_str());
  exit(3);
}

static int run_rank(const Args& a, int rank, int N, const string& dir) {
  const vector<long> counts = parse_list<long>(a.get("counts", "1024"));
  const int nblobs = (int)counts.size();
  vector<float> lr_mult = parse_list<float>(a.get("lr_mult", ""));
  vector<float> decay_mult = parse_list<float>(a.get("decay_mult", ""));
  lr_mult.resize(nblobs, 1.0f);
  decay_mult.resize(nblobs, 1.0f);
  vector<int64_t> counts64(counts.begin(), counts.end());
  const int policy = policy_id(a.get("lr_policy", "fixed"));
  const float base_lr = (float)a.getf("base_lr", 0.01);
  const float gamma = (float)a.getf("gamma", 0.1);
  const float power = (float)a.getf("power", 0.75);
  const int stepsize = (int)a.geti("stepsize", 1);
  const int max_iter = (int)a.geti("max_iter", 1000);
  const vector<int> stepvalues = parse_list<int>(a.get("stepvalue", ""));
  const float momentum = (float)a.getf("momentum", 0.9);
  const float wd = (float)a.getf("weight_decay", 0.0005);
  const int iters = (int)a.geti("iters", 3);
  const uint64_t seed = (uint64_t)a.geti("seed", 1);
  const float w_amp = (float)a.getf("w_amp", 0.05);
  const float g_amp = (float)a.getf("g_amp", 0.01);
  const bool bf16 = a.geti("bf16", 0) != 0;
  const bool dump = a.geti("dump", 0) != 0;
  const bool timing = a.geti("time", 0) != 0;

  // --- cluster bring-up, in the order CaffeNet.cpp:253-272,456-480 uses ---
  vector<shared_ptr<caffe::SocketChannel> > channels(N);
  for (int i = 0; i < N; ++i)
    if (i != rank) channels[i].reset(new caffe::SocketChannel());
  caffe::SocketAdapter adapter(&channels);
  {
    // The reference publishes gethostname():port (socket.hpp:28-37); the
    // container hostname may not resolve, so the driver (playing the Spark
    // driver's role of moving address strings) rewrites the host part.
    char buf[64];
    snprintf(buf, sizeof(buf), "127.0.0.1:%d", (int)adapter.port);
    char name[64];
    snprintf(name, sizeof(name), "/addr_%d", rank);
    write_file_atomic(dir + name, buf);
  }
  vector<string> addrs(N);
  for (int i = 0; i < N; ++i) {
    if (i == rank) continue;
    char name[64];
    snprintf(name, sizeof(name), "/addr_%d", i);
    addrs[i] = wait_read_file(dir + name);
  }
  for (int i = 0; i < N; ++i) {
    if (i == rank) continue;
    if (!channels[i]->Connect(addrs[i])) {
      fprintf(stderr, "rank %d: connect to %s failed\n", rank, addrs[i].c_str());
      return 4;
    }
  }

  // --- solver with the requested learnable blob layout ---
  shared_ptr<caffe::Solver<float> > solver(new caffe::Solver<float>());
  uint64_t P = 0;
  for (int k = 0; k < nblobs; ++k) {
    solver->net()->add_param((int)counts[k]);
    P += counts[k];
  }
  {
    // initial weights: identical on all ranks, then copied into data_ by the
    // CPUParams ctor (parallel_cpu.cpp:77-83)
    vector<float> w0(P ? P : 1);
    cos_oracle_fill(P, w0.data(), seed, 0, w_amp);
    uint64_t o = 0;
    for (int k = 0; k < nblobs; ++k) {
      memcpy(solver->net()->learnable_params()[k]->mutable_cpu_data(),
             w0.data() + o, counts[k] * sizeof(float));
      o += counts[k];
    }
  }
  caffe::Caffe::set_solver_count(N);  // CaffeNet.cpp:625
  caffe::SocketSyncCPU<float> sync(solver, channels, rank);
  if (sync.size() != (P ? P : 1)) return 5;
  float* data = sync.data();
  float* diff = sync.diff();
  vector<float> hist(sync.size(), 0.f);  // SGDSolver::PreSolve zero history

  uint64_t own_offs, own_size;
  cos_oracle_chunk(sync.size(), N, rank, &own_offs, &own_size);

  FILE* dumpf = NULL;
  if (dump) {
    char name[64];
    snprintf(name, sizeof(name), "/rank_%d.bin", rank);
    dumpf = fopen((dir + name).c_str(), "wb");
  }
  vector<double> t_sync(iters), t_update(iters);
  int current_step = 0;
  sync.sync(false);  // CaffeProcessor.sync(): control barrier before feeding
  for (int t = 0; t < iters; ++t) {
    memset(diff, 0, sync.size() * sizeof(float));  // ClearParamDiffs
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    solver->fire_on_start();
    std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
    // ForwardBackward stand-in: synthetic local gradient of (rank, iter)
    cos_oracle_fill(P, diff, seed, (uint64_t)(t + 1) * 4096 + (uint64_t)rank, g_amp);
    if (bf16) cos_oracle_round_bf16(P, diff);
    std::chrono::steady_clock::time_point t2 = std::chrono::steady_clock::now();
    solver->fire_on_gradients_ready();
    std::chrono::steady_clock::time_point t3 = std::chrono::steady_clock::now();
    float rate = cos_oracle_learning_rate(policy, base_lr, gamma, power, stepsize,
                                          stepvalues.empty() ? NULL : stepvalues.data(),
                                          (int)stepvalues.size(), max_iter,
                                          solver->iter(), &current_step);
    cos_oracle_apply_update(0, P, data, diff, hist.data(), nblobs, counts64.data(),
                            lr_mult.data(), decay_mult.data(), rate, momentum, wd);
    solver->advance();
    std::chrono::steady_clock::time_point t4 = std::chrono::steady_clock::now();
    t_sync[t] = std::chrono::duration<double, std::milli>(t1 - t0).count() +
                std::chrono::duration<double, std::milli>(t3 - t2).count();
    t_update[t] = std::chrono::duration<double, std::milli>(t4 - t3).count();
    if (dumpf) {
      fwrite(data + own_offs, sizeof(float), own_size, dumpf);
      fwrite(hist.data() + own_offs, sizeof(float), own_size, dumpf);
    }
  }
  // trailing on_start: every rank ends with the consistent weights
  solver->fire_on_start();
  if (dumpf) {
    fwrite(data, sizeof(float), P, dumpf);
    fclose(dumpf);
  }
  if (timing) {
    char name[64];
    snprintf(name, sizeof(name), "/time_%d.txt", rank);
    FILE* f = fopen((dir + name).c_str(), "w");
    for (int t = 0; t < iters; ++t) fprintf(f, "%.6f %.6f\n", t_sync[t], t_update[t]);
    fclose(f);
  }
  sync.sync(false);  // keep sockets alive until everyone is done
  return 0;
}

static double median(vector<double> v) {
  if (v.empty()) return 0;
  std::sort(v.begin(), v.end());
  return v[v.size() / 2];
}

int main(int argc, char** argv) {
  Args a;
  for (int i = 1; i < argc; ++i) {
    string s(argv[i]);
    size_t eq = s.find('=');
    if (s.compare(0, 2, "--") != 0 || eq == string::npos) {
      fprintf(stderr, "usage: ref_sync --key=value ... (ranks, counts, lr_mult, decay_mult, "
                      "lr_policy, base_lr, gamma, power, stepsize, stepvalue, max_iter, momentum, "
                      "weight_decay, iters, seed, w_amp, g_amp, bf16, dump, time, dir)\n");
      return 2;
    }
    a.kv[s.substr(2, eq - 2)] = s.substr(eq + 1);
  }
  const int N = (int)a.geti("ranks", 2);
  string dir = a.get("dir", "");
  if (dir.empty()) {
    char tmpl[] = "/tmp/cos_ref_XXXXXX";
    dir = mkdtemp(tmpl);
  } else {
    mkdir(dir.c_str(), 0777);
  }
  // stale rendezvous files from a previous run in the same dir
  for (int r = 0; r < N; ++r) {
    char name[64];
    snprintf(name, sizeof(name), "/addr_%d", r);
    unlink((dir + name).c_str());
  }
  vector<pid_t> pids(N);
  for (int r = 0; r < N; ++r) {
    pid_t p = fork();
    if (p == 0) {
      int rc = run_rank(a, r, N, dir);
      fflush(NULL);
      _exit(rc);  // listener/receiver threads never join (socket.cpp:203)
    }
    pids[r] = p;
  }
  int bad = 0;
  for (int r = 0; r < N; ++r) {
    int st = 0;
    waitpid(pids[r], &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1;
  }
  if (bad) {
    fprintf(stderr, "ref_sync: a rank failed\n");
    return 1;
  }
  if (a.geti("time", 0)) {
    const int iters = (int)a.geti("iters", 3);
    vector<double> it_total(iters, 0.0), it_sync(iters, 0.0);
    for (int r = 0; r < N; ++r) {
      char name[64];
      snprintf(name, sizeof(name), "/time_%d.txt", r);
      FILE* f = fopen((dir + name).c_str(), "r");
      for (int t = 0; t < iters; ++t) {
        double s = 0, u = 0;
        if (fscanf(f, "%lf %lf", &s, &u) != 2) break;
        it_total[t] = std::max(it_total[t], s + u);  // max over ranks
        it_sync[t] = std::max(it_sync[t], s);
      }
      fclose(f);
    }
    // drop iteration 0 (connection warm-up), median of the rest
    vector<double> tt(it_total.begin() + (iters > 1 ? 1 : 0), it_total.end());
    vector<double> ts(it_sync.begin() + (iters > 1 ? 1 : 0), it_sync.end());
    printf("{\"ranks\": %d, \"iters\": %d, \"ms_per_iter_median\": %.6f, "
           "\"ms_per_iter_min\": %.6f, \"ms_sync_median\": %.6f, \"cores\": %ld}\n",
           N, iters, median(tt), *std::min_element(tt.begin(), tt.end()), median(ts),
           sysconf(_SC_NPROCESSORS_ONLN));
  }
  printf("dir=%s\n", dir.c_str());
  return 0;
}
