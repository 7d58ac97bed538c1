// TEST INFRASTRUCTURE.  Does the reference's current engine run two processes in one simulation (what cProcess::send,
// inst/include/microsimulation.h:365-384, would need)?  It does not: Sim::create_process pushes an A_Init action with a
// null event (src/ssim.cc:108-113), and Action::operator< reads event->priority of both actions when their times are
// equal (src/ssim.cc:60-62) — the second create_process, also at time 0, dereferences null.  (inst/include/ssim.cc, the
// stale copy whose operator< ignores priority, is not the one the package builds: src/Makevars:5.)
//   g++ -std=gnu++17 -Wno-deprecated -Ioracle/shim -I/root/reference/inst/include tests/manual/probe_two_processes.cc \
//       /root/reference/src/ssim.cc && ./a.out          -> "one created", then SIGSEGV
// tests/test_oracle_ref.py runs it where /root/reference exists.
#include <stdio.h>
#include <siena/ssim.h>
using namespace ssim;
struct P : public ProcessWithPId {
  const char *name;
  P(const char *n) : name(n) {}
  void initialize() { printf("%s init at %g\n", name, Sim::clock()); }
  void process_event(const Event *) {}
};
int main() {
  P a("a"), b("b");
  Sim::create_process(&a);
  printf("one created\n"); fflush(stdout);
  Sim::create_process(&b);
  printf("two created\n"); fflush(stdout);
  Sim::run_simulation();
  Sim::clear();
  return 0;
}
