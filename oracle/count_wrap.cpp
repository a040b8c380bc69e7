// TEST / MEASUREMENT INFRASTRUCTURE: counts the NucCruc heterodimer evaluations of a reference run.
//
// oracle/_ref/tntblast_counted is the reference linked with
//     -Wl,--wrap=_ZN7NucCruc26approximate_tm_heterodimerEv
// so that every call bind_oligo.cpp makes to NucCruc::approximate_tm_heterodimer (bind_oligo.cpp:595,
// :1298, and the per-hit primer-dimer call of tntblast_local.cpp:683) passes through this counter
// before it reaches the unmodified function.  The count is printed on stderr when the process
// ends; bench.py reports it next to the CPU timing (which is taken with the plain binary).
#include <atomic>
#include <cstdio>
#include <cstdlib>

class NucCruc;
extern "C" float __real__ZN7NucCruc26approximate_tm_heterodimerEv(NucCruc *self);

static std::atomic<unsigned long long> g_calls(0);

static void report() { fprintf(stderr, "[tntref] approximate_tm_heterodimer calls: %llu\n", g_calls.load()); }

struct Registrar { Registrar() { atexit(report); } };
static Registrar g_registrar;

extern "C" float __wrap__ZN7NucCruc26approximate_tm_heterodimerEv(NucCruc *self)
{
	g_calls.fetch_add(1, std::memory_order_relaxed);
	return __real__ZN7NucCruc26approximate_tm_heterodimerEv(self);
}
