// workers_check.cpp — test program for crdmodel_b200/host/crd_workers.hpp: ranks that block on a process-shared barrier
// (like the drivers' ranks do on each other) while one of them misbehaves.  usage: workers_check <mode> <nranks>
//   ok      every rank passes the barrier and returns 0                     -> exit 0
//   fail    rank 2 returns 3 before the barrier, the others wait forever    -> exit 1, promptly
//   crash   rank 1 aborts before the barrier                                -> exit 1, promptly
//   rank0   rank 0 returns 5 before the barrier                             -> exit 5, workers killed
#include <pthread.h>
#include <sys/mman.h>

#include <cstdlib>
#include <cstring>

#include "crd_workers.hpp"

int main(int argc, char **argv) {
  if (argc != 3) return 2;
  const char *mode = argv[1];
  const int nranks = std::atoi(argv[2]);
  pthread_barrier_t *bar = (pthread_barrier_t *)mmap(NULL, sizeof(pthread_barrier_t), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (bar == MAP_FAILED) return 2;
  pthread_barrierattr_t ba;
  pthread_barrierattr_init(&ba);
  pthread_barrierattr_setpshared(&ba, PTHREAD_PROCESS_SHARED);
  pthread_barrier_init(bar, &ba, nranks);
  return crd::run_ranks(nranks, [&](int r) {
    if (!std::strcmp(mode, "fail") && r == 2) return 3;
    if (!std::strcmp(mode, "crash") && r == 1) std::abort();
    if (!std::strcmp(mode, "rank0") && r == 0) { usleep(100000); return 5; }
    pthread_barrier_wait(bar);
    return 0;
  });
}
