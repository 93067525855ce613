/* Test infrastructure (not product code): exhaustive check of the division-free sequence gram_reforder_kernel
 * (grav1synth_b200/csrc/g1s_gram_strict.cu) uses for the reference's `(buffer[i] * buffer[j]) / (255 * 255)`
 * (libaom noise_model.c add_block_observations; av1-grain diff/solver.rs, reached from src/main.rs:442):
 *     q0 = p*y;  r = fma(-q0, 65025, p);  q = fma(r, y, q0),   y = RN(1/65025)
 * equals p / 65025.0 (IEEE, round to nearest) for every integer |p| <= 1020^2.  Prints the number of mismatches.
 *   gcc -O2 -ffp-contract=off -o div65025_check div65025_check.c -lm && ./div65025_check */
#include <math.h>
#include <stdio.h>
int main(void) {
  const double y = 1.0 / 65025.0;
  long bad = 0, single = 0;
  for (long p = -1040400; p <= 1040400; ++p) {
    const double pd = (double)p, q0 = pd * y, r = fma(-q0, 65025.0, pd), q = fma(r, y, q0), t = pd / 65025.0;
    bad += q != t;
    single += q0 != t;
  }
  printf("mismatches=%ld (one multiplication alone would miss %ld)\n", bad, single);
  return bad != 0;
}
