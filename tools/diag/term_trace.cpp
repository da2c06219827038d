// Development aid: LD_PRELOAD this to get a native backtrace when std::terminate() fires at interpreter exit.
//   g++ -shared -fPIC -O1 -o gpurun_out/libterm_trace.so tools/diag/term_trace.cpp
#include <execinfo.h>
#include <unistd.h>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <sys/syscall.h>

static void dump(const char* why) {
  void* frames[96];
  const int n = backtrace(frames, 96);
  dprintf(2, "=== term_trace: %s (tid %ld, pid %d) ===\n", why, (long)syscall(SYS_gettid), (int)getpid());
  backtrace_symbols_fd(frames, n, 2);
  dprintf(2, "=== end ===\n");
}
static void on_terminate() {
  dump("std::terminate");
  _exit(134);
}
static void on_abort(int) {
  dump("SIGABRT");
  _exit(134);
}
__attribute__((constructor)) static void install() {
  std::set_terminate(on_terminate);
  signal(SIGABRT, on_abort);
}
