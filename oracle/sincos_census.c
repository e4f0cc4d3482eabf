/* Census of last-bit disagreements between three sin/cos routines -- TEST INFRASTRUCTURE ONLY.
 *
 *   det   aur_ppo_b200/csrc/det_sincos.h      what the CUDA env kernels evaluate (and oracle/envs.c in TRIG_DET mode)
 *   libm  the host's glibc sin / cos          what gym's cartpole.py / pendulum.py / acrobot.py call (math.sin, np.sin)
 *   cr    (double) sinq / cosq of libquadmath the checker's INDEPENDENT reference: 113-bit evaluation rounded once to
 *                                             double = the correctly rounded result (a double-rounding miss needs ~50
 *                                             equal bits after the 53rd: probability ~1e-15 per point)
 * over N random doubles per reachable range of the five env ids.
 *   gcc -O2 -fopenmp -ffp-contract=off oracle/sincos_census.c -lquadmath -lm -o oracle/_ref/sincos_census
 *   oracle/_ref/sincos_census 100000000
 */
#include <math.h>
#include <quadmath.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../aur_ppo_b200/csrc/det_sincos.h"

static inline uint64_t splitmix(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

int main(int argc, char** argv) {
  const long long N = argc > 1 ? atoll(argv[1]) : 10000000LL;
  const struct { const char* name; double lo, hi; } R[] = {
      {"CartPole theta (|theta| <= 2 x 12 deg threshold)", -0.42, 0.42},
      {"Pendulum theta (unwrapped, 200 steps at |thdot| <= 8)", -85.0, 85.0},
      {"Acrobot joint angles and sums ([-2 pi, 2 pi])", -6.3, 6.3},
      {"MountainCar 3 * position ([-3.6, 1.8])", -3.6, 1.8},
  };
  printf("| range | points | det != cr (sin) | det != cr (cos) | libm != cr (sin) | libm != cr (cos) | det != libm (sin or cos) |\n");
  printf("|---|---:|---:|---:|---:|---:|---:|\n");
  for (int r = 0; r < 4; ++r) {
    long long ds = 0, dc = 0, ls = 0, lc = 0, dl = 0;
#pragma omp parallel for reduction(+ : ds, dc, ls, lc, dl) schedule(static)
    for (long long i = 0; i < N; ++i) {
      uint64_t st = 0x1234567ULL * (r + 1) + (uint64_t)i * 0xD1342543DE82EF95ULL;
      const double u = (double)(splitmix(&st) >> 11) * (1.0 / 9007199254740992.0);
      const double x = R[r].lo + (R[r].hi - R[r].lo) * u;
      double s1, c1;
      aur_sincos(x, &s1, &c1);
      const double s2 = sin(x), c2 = cos(x);
      const double s3 = (double)sinq((__float128)x), c3 = (double)cosq((__float128)x);
      ds += s1 != s3; dc += c1 != c3; ls += s2 != s3; lc += c2 != c3; dl += (s1 != s2) || (c1 != c2);
    }
    printf("| %s | %lld | %lld | %lld | %lld | %lld | %lld |\n", R[r].name, N, ds, dc, ls, lc, dl);
    fflush(stdout);
  }
  return 0;
}
