// oracle/shim/glog/logging.h -- TEST INFRASTRUCTURE, not product code.
//
// Minimal stand-in for <glog/logging.h> so that the reference's hot-path
// sources (which include it from core/pbrt.h:61) compile straight from
// /root/reference with plain g++ and no cmake-generated headers.  Only the
// macros the reference actually uses are provided: LOG(sev), VLOG(n),
// CHECK*/DCHECK*, the FLAGS_* variables main/pbrt.cpp pokes and
// google::InitGoogleLogging.  Semantics kept: FATAL and failed CHECKs abort,
// WARNING and above go to stderr, INFO/VLOG are discarded unless
// FLAGS_logtostderr / FLAGS_v ask for them.  Nothing here touches arithmetic.
#ifndef B200PT_ORACLE_GLOG_SHIM_H
#define B200PT_ORACLE_GLOG_SHIM_H

#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

extern int FLAGS_stderrthreshold;
extern int FLAGS_minloglevel;
extern int FLAGS_v;
extern bool FLAGS_logtostderr;
extern std::string FLAGS_log_dir;

namespace google {
enum { GLOG_INFO = 0, GLOG_WARNING = 1, GLOG_ERROR = 2, GLOG_FATAL = 3 };
inline void InitGoogleLogging(const char *) {}

class ShimLogMessage {
  public:
    ShimLogMessage(const char *file, int line, int sev) : sev_(sev) {
        static const char *names[] = {"I", "W", "E", "F"};
        os_ << names[sev] << " " << file << ":" << line << "] ";
    }
    ~ShimLogMessage() {
        bool show = sev_ >= GLOG_WARNING || FLAGS_logtostderr;
        if (sev_ < FLAGS_minloglevel && sev_ != GLOG_FATAL) show = false;
        if (show) std::cerr << os_.str() << std::endl;
        if (sev_ == GLOG_FATAL) std::abort();
    }
    std::ostream &stream() { return os_; }

  private:
    int sev_;
    std::ostringstream os_;
};

// Swallows a stream expression so "cond ? (void)0 : Voidify() & stream" types.
struct ShimVoidify {
    void operator&(std::ostream &) {}
};

template <typename T>
inline T &ShimCheckNotNull(const char *file, int line, const char *expr, T &t) {
    if (t == nullptr) {
        ShimLogMessage(file, line, GLOG_FATAL).stream() << "Check failed: " << expr;
    }
    return t;
}
}  // namespace google

#define B200PT_SHIM_LOG_INFO ::google::GLOG_INFO
#define B200PT_SHIM_LOG_WARNING ::google::GLOG_WARNING
#define B200PT_SHIM_LOG_ERROR ::google::GLOG_ERROR
#define B200PT_SHIM_LOG_FATAL ::google::GLOG_FATAL

#define LOG(sev) \
    ::google::ShimLogMessage(__FILE__, __LINE__, B200PT_SHIM_LOG_##sev).stream()

#define B200PT_SHIM_COND_LOG(cond, sev)      \
    !(cond) ? (void)0                        \
            : ::google::ShimVoidify() & LOG(sev)

#define VLOG(n) B200PT_SHIM_COND_LOG(FLAGS_v >= (n), INFO)
#define LOG_IF(sev, cond) B200PT_SHIM_COND_LOG(cond, sev)

#define CHECK(cond) \
    B200PT_SHIM_COND_LOG(!(cond), FATAL) << "Check failed: " #cond " "

#define B200PT_SHIM_CHECK_OP(a, b, op)                                      \
    B200PT_SHIM_COND_LOG(!((a)op(b)), FATAL)                                \
        << "Check failed: " #a " " #op " " #b " (" << (a) << " vs. " << (b) \
        << ") "

#define CHECK_EQ(a, b) B200PT_SHIM_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) B200PT_SHIM_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) B200PT_SHIM_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) B200PT_SHIM_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) B200PT_SHIM_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) B200PT_SHIM_CHECK_OP(a, b, >=)
#define CHECK_NOTNULL(p) \
    ::google::ShimCheckNotNull(__FILE__, __LINE__, "'" #p "' Must be non NULL", (p))

#ifdef NDEBUG
#define B200PT_SHIM_DEAD(expr) \
    true ? (void)0 : ::google::ShimVoidify() & LOG(INFO) << (expr)
#define DCHECK(cond) B200PT_SHIM_DEAD(cond)
#define DCHECK_EQ(a, b) B200PT_SHIM_DEAD((a) == (b))
#define DCHECK_NE(a, b) B200PT_SHIM_DEAD((a) != (b))
#define DCHECK_LT(a, b) B200PT_SHIM_DEAD((a) < (b))
#define DCHECK_LE(a, b) B200PT_SHIM_DEAD((a) <= (b))
#define DCHECK_GT(a, b) B200PT_SHIM_DEAD((a) > (b))
#define DCHECK_GE(a, b) B200PT_SHIM_DEAD((a) >= (b))
#else
#define DCHECK(cond) CHECK(cond)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) CHECK_NE(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_LE(a, b) CHECK_LE(a, b)
#define DCHECK_GT(a, b) CHECK_GT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#endif

#endif  // B200PT_ORACLE_GLOG_SHIM_H
