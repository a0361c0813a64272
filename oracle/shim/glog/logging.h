// Oracle build shim (test infrastructure): the handful of glog macros the
// reference's socket.cpp / socket_sync_cpu.cpp / parallel_cpu.cpp use.
#ifndef COS_SHIM_GLOG_LOGGING_H_
#define COS_SHIM_GLOG_LOGGING_H_
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace cos_shim {
struct LogLine {
  bool fatal, show;
  std::ostringstream os;
  LogLine(bool f, bool s) : fatal(f), show(s) {}
  ~LogLine() {
    if (show || fatal) std::cerr << os.str() << std::endl;
    if (fatal) std::abort();
  }
  template <typename T> LogLine& operator<<(const T& v) { os << v; return *this; }
  LogLine& operator<<(std::ostream& (*m)(std::ostream&)) { os << m; return *this; }
};
struct Voidify { void operator&(const LogLine&) {} };
inline bool verbose() { static bool v = std::getenv("COS_REF_VERBOSE") != 0; return v; }
template <typename T> T* check_notnull(T* p) { if (!p) std::abort(); return p; }
}  // namespace cos_shim
#define COS_SEV_INFO false, ::cos_shim::verbose()
#define COS_SEV_WARNING false, ::cos_shim::verbose()
#define COS_SEV_ERROR false, true
#define COS_SEV_FATAL true, true
#define LOG(sev) ::cos_shim::LogLine(COS_SEV_##sev)
#define DLOG(sev) LOG(sev)
#define LOG_IF(sev, c) !(c) ? (void)0 : ::cos_shim::Voidify() & LOG(sev)
#define CHECK(c) (c) ? (void)0 : ::cos_shim::Voidify() & ::cos_shim::LogLine(true, true) << "Check failed: " #c " "
#define CHECK_OP_(a, b, op) CHECK((a) op (b))
#define CHECK_EQ(a, b) CHECK_OP_(a, b, ==)
#define CHECK_NE(a, b) CHECK_OP_(a, b, !=)
#define CHECK_GE(a, b) CHECK_OP_(a, b, >=)
#define CHECK_GT(a, b) CHECK_OP_(a, b, >)
#define CHECK_LE(a, b) CHECK_OP_(a, b, <=)
#define CHECK_LT(a, b) CHECK_OP_(a, b, <)
#define CHECK_NOTNULL(p) ::cos_shim::check_notnull(p)
#endif
