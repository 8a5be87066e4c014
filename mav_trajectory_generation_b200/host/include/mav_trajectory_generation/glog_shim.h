// glog_shim.h -- CHECK / LOG macros with glog's spelling and abort-on-failure semantics, used
// by the host mirror exactly where the reference uses them (argument errors abort the process,
// e.g. reference impl/polynomial_optimization_linear_impl.h:60,76,289,297).  If the real glog
// is available it is used instead.
#ifndef MAV_TRAJECTORY_GENERATION_GLOG_SHIM_H_
#define MAV_TRAJECTORY_GENERATION_GLOG_SHIM_H_

#if defined(__has_include)
#if __has_include(<glog/logging.h>) && !defined(MTG_FORCE_GLOG_SHIM)
#define MTG_HAVE_GLOG 1
#endif
#endif

#ifdef MTG_HAVE_GLOG
#include <glog/logging.h>
#else
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace mtg_log {
class Message {
 public:
  Message(const char* file, int line, const char* level, bool fatal) : fatal_(fatal) {
    stream_ << "[" << level << " " << file << ":" << line << "] ";
  }
  ~Message() {
    std::cerr << stream_.str() << std::endl;
    if (fatal_) std::abort();
  }
  std::ostream& stream() { return stream_; }

 private:
  std::ostringstream stream_;
  bool fatal_;
};
struct Voidify {
  void operator&(std::ostream&) {}
};
template <typename T>
T* CheckNotNull(const char* file, int line, const char* expr, T* p) {
  if (p == nullptr) Message(file, line, "FATAL", true).stream() << "'" << expr << "' Must be non NULL";
  return p;
}
}  // namespace mtg_log

#define MTG_LOG_INFO ::mtg_log::Message(__FILE__, __LINE__, "INFO", false).stream()
#define MTG_LOG_WARNING ::mtg_log::Message(__FILE__, __LINE__, "WARNING", false).stream()
#define MTG_LOG_ERROR ::mtg_log::Message(__FILE__, __LINE__, "ERROR", false).stream()
#define MTG_LOG_FATAL ::mtg_log::Message(__FILE__, __LINE__, "FATAL", true).stream()
#define LOG(severity) MTG_LOG_##severity
#define DLOG(severity) \
  true ? (void)0 : ::mtg_log::Voidify() & MTG_LOG_##severity
#define VLOG(level) true ? (void)0 : ::mtg_log::Voidify() & MTG_LOG_INFO
#define CHECK(cond) \
  (cond) ? (void)0 : ::mtg_log::Voidify() & MTG_LOG_FATAL << "Check failed: " #cond " "
#define MTG_CHECK_OP(a, b, op) CHECK((a)op(b)) << "(" << (a) << " vs. " << (b) << ") "
#define CHECK_EQ(a, b) MTG_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) MTG_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) MTG_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) MTG_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) MTG_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) MTG_CHECK_OP(a, b, >=)
#define CHECK_NOTNULL(p) ::mtg_log::CheckNotNull(__FILE__, __LINE__, #p, (p))
#endif  // MTG_HAVE_GLOG
#endif  // MAV_TRAJECTORY_GENERATION_GLOG_SHIM_H_
