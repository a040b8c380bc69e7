// CPU study for the pre-alignment score bound (not part of the product path).
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

// Lagrangian lower bound of the DP score of a gapless core alignment that reaches min_tm
static double smin_gapless(const Thermo &th, const OligoStrand &os, const int32_t *rows, float min_tm, double *best_lambda)
{
	const int L = os.len;
	const double K = (double)min_tm - 0.05 + 273.15;
	const double limit = K*(double)os.r_log_ct;
	const double c_s = -K*(double)th.salt*(double)th.log_na;
	const double at_g = (double)th.at_H - K*(double)th.at_S;
	const double INF = 1e300;
	auto code_of = [&](int x, int t) { return (int)th.bbp[os.seq[x]*NB + t]; };
	auto is_at = [&](int c) { return c == 7*bA + bT || c == 7*bT + bA; };
	const int MM = L + 1;
	double best_bound = 0;
	for (int li = 0; li <= 400; ++li) {
		const double lambda = 2000.0 + 50.0*li; // score units per kcal
		// f[x][t][m]: min of score + lambda*G_K over alignments ending at (x,t) with open mismatch run m
		std::vector<double> f((size_t)L*4*MM, INF);
		double best = INF;
		for (int x = 0; x < L; ++x) {
			// extend from x-1
			for (int t = 0; t < 4; ++t) {
				const int c = code_of(x, t);
				if (th.wc[c]) {
					double &s = f[(size_t)(x*4 + t)*MM];
					s = std::min(s, lambda*((double)th.init_H - K*(double)th.init_S + (is_at(c) ? at_g : 0.0)));
				}
			}
			if (x + 1 >= L) break;
			for (int t1 = 0; t1 < 4; ++t1) {
				const int last = code_of(x, t1);
				for (int m = 0; m < MM; ++m) {
					const double base = f[(size_t)(x*4 + t1)*MM + m];
					if (base >= INF) continue;
					for (int t2 = 0; t2 < 4; ++t2) {
						const int cur = code_of(x + 1, t2);
						// DP gain of this stack: row index of oligo position x+1 ... rows are reversed:
						// row i pairs q[L-i]; the column order along the diagonal runs with i, i.e. against x.
						// Use the table of the later row in DP order = smaller x.  Handled by the caller via
						// a symmetric helper: gain(x, t1, x+1, t2)
						double v = base;
						// DP row for oligo position x (i = L - x), previous pair is (x+1, t2) in DP order
						const int i = L - x; // row of position x; its predecessor row i-1 is position x+1
						const int32_t *row = rows + (size_t)(i - 1)*72;
						const int p1 = row[t2*4 + t1]; // pt = target base of the previous DP column = t2, tb = t1
						v += -(double)p1*1.0; // score contribution is -P1 ... we minimise score => add gain
						// careful: score = sum of gains = sum(-p1); we minimise score + lambda*G
						v = base + (double)(-p1);
						double g = 0;
						if (th.wc[last] || th.wc[cur]) g += (double)th.H[last*NPAIR + cur] - K*(double)th.S[last*NPAIR + cur] + c_s;
						int m2;
						if (th.wc[cur]) {
							if (m > 1) g += -K*(double)th.loop_S[2*m] + c_s;
							m2 = 0;
						}
						else m2 = m + 1;
						v += lambda*g;
						if (m2 >= MM) continue;
						double &slot = f[(size_t)((x + 1)*4 + t2)*MM + m2];
						if (v < slot) slot = v;
					}
				}
			}
		}
		for (int x = 0; x < L; ++x)
			for (int t = 0; t < 4; ++t) {
				const int c = code_of(x, t);
				if (th.wc[c]) best = std::min(best, f[(size_t)(x*4 + t)*MM] + lambda*(is_at(c) ? at_g : 0.0));
			}
		// note: single-column alignments are included (score 0, G = init): harmless, they lower the bound
		const double bound = best - lambda*limit;
		if (bound > best_bound) { best_bound = bound; if (best_lambda) *best_lambda = lambda; }
	}
	return best_bound;
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
		double lam = 0;
		const double smin = smin_gapless(*th, os, rows.data(), min_tm, &lam);
		// perfect-match score
		const int Lt = L + 8;
		std::vector<uint8_t> tgt(Lt);
		long n_below = 0, n_gapless_eq = 0, n_g1_below = 0;
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
			if (s < smin) ++n_below;
			if (g1 < smin) ++n_g1_below;
			if (s == g1) ++n_gapless_eq;
		}
		// perfect match
		for (int j = 0; j < Lt; ++j) tgt[j] = rng() & 3;
		for (int i = 1; i <= L; ++i) tgt[i + 4 - 1] = 3 - os.seq[L - i];
		const int sperf = full_dp(rows.data(), p5, L, tgt.data(), Lt, nullptr);
		std::sort(sstar.begin(), sstar.end());
		printf("oligo %d: mincols %d smin %.0f (lambda %.0f) perfect %d | S* med %d p90 %d p99 %d | S*<smin %.3f  G1<smin %.3f  S*==G1 %.3f\n",
			o, mincols, smin, lam, sperf, sstar[nwin/2], sstar[nwin*9/10], sstar[nwin*99/100],
			(double)n_below/nwin, (double)n_g1_below/nwin, (double)n_gapless_eq/nwin);
	}
	return 0;
}
