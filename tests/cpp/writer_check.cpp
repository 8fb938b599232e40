#include "crd_writer.hpp"
#include <cmath>
#include <cstring>
// Test helper: crd::AsyncWriter must produce the bytes of the reference's fprintf(" %.16e") loop.
int main(int argc, char **argv) {
  if (argc != 4) return 2;
  const long n = 100003;
  std::vector<double> s(2 * n);
  for (long k = 0; k < 2 * n; ++k) s[k] = std::sin(0.37 * k) * std::pow(10.0, (k % 40) - 20) * ((k % 3) ? 1 : -1);
  s[5] = 0.0; s[7] = -0.0; s[9] = 1e-310; s[11] = 1e300;
  FILE *a = fopen(argv[1], "w"), *b = fopen(argv[2], "w");
  { crd::AsyncWriter w(a, b, true, n, 8); w.submit(s.data()); w.submit(s.data()); w.finish(); }
  fclose(a); fclose(b);
  FILE *r = fopen(argv[3], "w");
  for (int rep = 0; rep < 2; ++rep) { for (long k = 0; k < n; ++k) fprintf(r, " %.16e", s[2 * k]); fprintf(r, "\n"); }
  fclose(r);
  return 0;
}
