// CPU study for a pre-alignment score bound (not part of the product path):
// how the best local-alignment score of seeded random windows compares with a perfect match, and how often
// the optimum is gapless (the only case the minimum-columns bound of the lean tier speaks about).
//   g++ -O2 -I thermonucleotideblast_b200/csrc tools/prefilter_study.cpp thermonucleotideblast_b200/csrc/thermo.cpp -o /tmp/pfstudy
#include "thermo.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <random>

using namespace tnt;

static int full_dp(const int32_t *rows, const int32_t *p5, int L, const uint8_t *tgt, int Lt, bool *gapless_only)
{
	std::vector<int> M((L + 1)*(Lt + 1), 0), Iq((L + 1)*(Lt + 1), 0), It((L + 1)*(Lt + 1), 0);
	int best = 0;
	auto at = [&](std::vector<int> &v, int i, int j) -> int & { return v[i*(Lt + 1) + j]; };
	for (int i = 1; i <= L; ++i) {
		const int32_t *row = rows + (size_t)(i - 1)*72;
		for (int j = 1; j <= Lt; ++j) {
			const int tb = tgt[j - 1];
			const int pt = j == 1 ? 4 : tgt[j - 2];
			const int td = pt*4 + tb;
			const int d1 = at(M, i - 1, j - 1) - row[0 + td];
			const int d2 = at(Iq, i - 1, j - 1) - row[20 + td];
			const int d3 = at(It, i - 1, j - 1) - row[60 + tb];
			const int m = std::max(d1, std::max(d2, d3));
			const int qi = at(M, i, j - 1) - row[40 + td];
			const int qe = at(Iq, i, j - 1) - p5[td];
			const int ti = at(M, i - 1, j) - row[64 + tb];
			const int te = at(It, i - 1, j) - row[68];
			best = std::max(best, m);
			at(M, i, j) = std::max(m, 0);
			at(Iq, i, j) = std::max(std::max(qi, qe), 0);
			at(It, i, j) = std::max(std::max(ti, te), 0);
		}
	}
	(void)gapless_only;
	return best;
}

static int gapless_dp(const int32_t *rows, int L, const uint8_t *tgt, int Lt)
{
	std::vector<int> M((L + 1)*(Lt + 1), 0);
	int best = 0;
	for (int i = 1; i <= L; ++i) {
		const int32_t *row = rows + (size_t)(i - 1)*72;
		for (int j = 1; j <= Lt; ++j) {
			const int tb = tgt[j - 1];
			const int pt = j == 1 ? 4 : tgt[j - 2];
			const int m = M[(i - 1)*(Lt + 1) + j - 1] - row[pt*4 + tb];
			best = std::max(best, m);
			M[i*(Lt + 1) + j] = std::max(m, 0);
		}
	}
	return best;
}

int main(int argc, char **argv)
{
	const int L = argc > 1 ? atoi(argv[1]) : 20;
	const float min_tm = argc > 2 ? (float)atof(argv[2]) : 45.0f;
	const int nolig = argc > 3 ? atoi(argv[3]) : 10;
	const int nwin = 20000;
	Thermo *th = new Thermo;
	build_thermo(*th, 310.15f, 0.05f, false, false);
	std::mt19937 rng(12345);
	int32_t p5[20];
	build_p5_table(*th, p5);
	printf("pen: p5[0]=%d\n", p5[0]);
	for (int o = 0; o < nolig; ++o) {
		OligoStrand os;
		memset(&os, 0, sizeof(os));
		os.len = L;
		for (int x = 0; x < L; ++x) os.seq[x] = rng() & 3;
		os.r_log_ct = r_log_ct(9.0e-7f);
		std::vector<int32_t> rows((size_t)L*72);
		build_row_tables(*th, os, rows.data());
		const int mincols = lean_min_columns(*th, os, min_tm);
		// perfect-match score
		const int Lt = L + 8;
		std::vector<uint8_t> tgt(Lt);
		long n_gapless_eq = 0;
		std::vector<int> sstar(nwin), g1s(nwin);
		for (int w = 0; w < nwin; ++w) {
			for (int j = 0; j < Lt; ++j) tgt[j] = rng() & 3;
			// plant a 7-mer on the main diagonal j = i + 4
			const int a = 1 + (int)(rng() % (L - 6));
			for (int k = 0; k < 7; ++k) {
				const int i = a + k, j = i + 4;
				tgt[j - 1] = 3 - os.seq[L - i];
			}
			const int s = full_dp(rows.data(), p5, L, tgt.data(), Lt, nullptr);
			const int g1 = gapless_dp(rows.data(), L, tgt.data(), Lt);
			sstar[w] = s; g1s[w] = g1;
			if (s == g1) ++n_gapless_eq;
		}
		// perfect match
		for (int j = 0; j < Lt; ++j) tgt[j] = rng() & 3;
		for (int i = 1; i <= L; ++i) tgt[i + 4 - 1] = 3 - os.seq[L - i];
		const int sperf = full_dp(rows.data(), p5, L, tgt.data(), Lt, nullptr);
		std::sort(sstar.begin(), sstar.end());
		printf("oligo %d: min columns %d, perfect-match score %d | best score of seeded random windows: median %d p90 %d p99 %d | gapless optimum == optimum in %.3f of the windows\n",
			o, mincols, sperf, sstar[nwin/2], sstar[nwin*9/10], sstar[nwin*99/100], (double)n_gapless_eq/nwin);
	}
	return 0;
}
