// tntblast_gpu: the reference program with its search path on the B200 engine.
//
// TEST / INTEGRATION INFRASTRUCTURE (see tntb200_shim.cpp).  main() of the reference
// (tntblast.cpp:28-77, the non-MPI branch) with one line added in front of local_main(): the batch
// pass of the shim.  Everything else -- option parsing, readers, the work loop, per-hit hairpin /
// dimer temperatures, uniquify, sorting, the text output -- is the reference's own object code,
// linked unmodified; its amplicon()/padlock()/hybrid() calls resolve to the shim.
#include <cstdlib>
#include <iostream>

#ifdef _OPENMP
#include <omp.h>
#endif

int local_main(int argc, char *argv[]);      // tntblast_local.cpp:25
bool tntb200_prefetch(int argc, char *argv[]);
void tntb200_report();

// globals the reference objects expect from tntblast.cpp:19-20
int mpi_numtasks;
int mpi_rank;

int main(int argc, char *argv[])
{
	mpi_numtasks = 1;
	mpi_rank = 0;
#ifdef _OPENMP
	std::cout << "Running on local machine [" << omp_get_max_threads() << " thread(s)]" << std::endl;
#else
	std::cout << "Running on local machine (1 thread)" << std::endl;
#endif
	try {
		tntb200_prefetch(argc, argv);
	}
	catch (const char *error) {
		std::cerr << "Caught the error: " << error << std::endl;
		return EXIT_FAILURE;
	}
	const int ret = local_main(argc, argv);
	tntb200_report();
	return ret;
}
