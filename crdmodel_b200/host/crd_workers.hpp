// crd_workers.hpp — the drivers' ranks: one worker process per GPU, forked by rank 0 (no MPI launcher).
//
// The ranks wait for each other on a shared barrier and on each other's halo flags, so when one worker dies (a CUDA error,
// a signal) the rest would wait forever.  Rank 0 therefore watches its children: the first abnormal exit ends the whole
// job with status 1, a failure of rank 0 itself takes the workers down with it, and no worker outlives its parent.
#pragma once
#include <signal.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstring>

namespace crd {

constexpr int kMaxWorkers = 64;
inline pid_t g_kids[kMaxWorkers];
inline int g_nkids = 0;
inline volatile sig_atomic_t g_stopping = 0;   // rank 0 is taking the workers down itself

inline void kill_workers() {
  for (int i = 0; i < g_nkids; ++i) kill(g_kids[i], SIGKILL);
}

inline void on_sigchld(int) {   // async-signal-safe calls only
  const int saved = errno;
  int st = 0;
  pid_t pid;
  while ((pid = waitpid(-1, &st, WNOHANG)) > 0) {
    if ((!WIFEXITED(st) || WEXITSTATUS(st) != 0) && !g_stopping) {
      kill_workers();
      static const char msg[] = "\nWORKER_ERROR: a GPU worker process failed, stopping the other ranks\n\n";
      if (write(2, msg, sizeof msg - 1) < 0) {}
      _exit(1);
    }
  }
  errno = saved;
}

// body(rank) runs in nranks processes (rank 0 = the caller); returns 0 when every rank returned 0
template <class Body>
int run_ranks(int nranks, Body body) {
  if (nranks > kMaxWorkers) nranks = kMaxWorkers;
  struct sigaction sa;
  std::memset(&sa, 0, sizeof sa);
  sa.sa_handler = on_sigchld;
  sa.sa_flags = SA_RESTART | SA_NOCLDSTOP;
  sigemptyset(&sa.sa_mask);
  sigaction(SIGCHLD, &sa, NULL);
  const pid_t parent = getpid();
  for (int r = 1; r < nranks; ++r) {
    fflush(NULL);
    pid_t pid = fork();
    if (pid < 0) { perror("fork"); kill_workers(); return 1; }
    if (pid == 0) {
      signal(SIGCHLD, SIG_DFL);
      prctl(PR_SET_PDEATHSIG, SIGKILL);        // never outlive the parent
      if (getppid() != parent) _exit(1);       // it died before the line above
      _exit(body(r));
    }
    g_kids[g_nkids++] = pid;
  }
  int rc = body(0);
  if (rc != 0) { g_stopping = 1; kill_workers(); }   // rank 0 failed: the others may be waiting for it
  for (int i = 0; i < g_nkids; ++i) {
    int st = 0;
    pid_t w;
    while ((w = waitpid(g_kids[i], &st, 0)) < 0 && errno == EINTR) {}
    if (w < 0) continue;                       // ECHILD: already reaped by the handler after a clean exit
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = rc ? rc : 1;
  }
  return rc;
}

}  // namespace crd
