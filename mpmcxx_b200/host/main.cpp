// mpmcxx-b200 INPUT [-P trotter] — the reference's command line (src/main.cpp, src/args_etc.h:216-292) over the host mirror:
// reads the input file, runs the Markov chain on the GPU engine, prints the running energy every corrtime.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mpmc_host.h"

using namespace mpmc_host;

int main(int argc, char **argv) {
	const char *input = nullptr;
	int P = 0;
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "-P") && i + 1 < argc) P = atoi(argv[++i]);
		else if (argv[i][0] != '-') input = argv[i];
	}
	if (!input) { fprintf(stderr, "usage: %s INPUT [-P trotter]\n", argv[0]); return 1; }
	try {
		SimulationControl sc((char *)input, P);
		sc.initializeSimulationObjects();
		std::vector<System::step_record> rec;
		sc.runSimulation(&rec);
		const unsigned ct = sc.sys.corrtime ? sc.sys.corrtime : 1;
		int acc = 0;
		for (size_t i = 0; i < rec.size(); i++) {
			acc += rec[i].accepted;
			if ((i + 1) % ct == 0 || i + 1 == rec.size())
				printf("step %zu  move %d  E_trial %.9f  BF %.6g  acceptance %.4f\n", i + 1, rec[i].movetype, rec[i].final_energy, rec[i].boltzmann_factor, (double)acc / (i + 1));
		}
		const System::observables_t &o = *sc.sys.observables;
		printf("final: energy %.9f  rd %.9f  coulombic %.9f  polarization %.9f  kinetic %.9f\n", o.energy, o.rd_energy, o.coulombic_energy, o.polarization_energy, o.kinetic_energy);
	} catch (int e) {
		fprintf(stderr, "MPMC exiting with error code: %d.\n", e);
		return 1;
	}
	return 0;
}
