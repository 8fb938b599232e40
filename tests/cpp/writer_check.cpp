#include "crd_writer.hpp"
#include <cmath>
#include <cstring>
// Test helper: crd::AsyncWriter must produce the bytes of the reference's fprintf(" %.16e") loop, take its data from the
// storage the producer names (no copy), and hand every buffer back exactly once, in order.
int main(int argc, char **argv) {
  if (argc != 4) return 2;
  const long n = 100003;
  std::vector<double> s(2 * n), v0(n), v1(n);
  for (long k = 0; k < 2 * n; ++k) s[k] = std::sin(0.37 * k) * std::pow(10.0, (k % 40) - 20) * ((k % 3) ? 1 : -1);
  s[5] = 0.0; s[7] = -0.0; s[9] = 1e-310; s[11] = 1e300; s[10] = -0.0; s[12] = 5e-324;
  for (long k = 0; k < n; ++k) { v0[k] = s[2 * k]; v1[k] = s[2 * k + 1]; }
  FILE *a = fopen(argv[1], "w"), *b = fopen(argv[2], "w");
  int released = 0, order_ok = 1;
  {
    crd::AsyncWriter w(a, b, true, n, 8);
    for (int rep = 0; rep < 3; ++rep)
      w.submit([&, rep] {
        crd::OutputView v;
        v.v0 = v0.data(); v.v1 = v1.data();
        v.release = [&, rep] { if (released != rep) order_ok = 0; ++released; };
        return v;
      });
    w.finish();
    if (w.failed()) return 3;
  }
  fclose(a); fclose(b);
  if (released != 3 || !order_ok) return 4;
  FILE *r = fopen(argv[3], "w");
  for (int rep = 0; rep < 3; ++rep) { for (long k = 0; k < n; ++k) fprintf(r, " %.16e", s[2 * k]); fprintf(r, "\n"); }
  fclose(r);
  return 0;
}
