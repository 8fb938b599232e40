// crd_steady.hpp — steady state (Zs, Ys) of the Goldbeter kinetics for spatially constant beta.
//
// The reference obtains it from an external script: popen("SolveGoldbeterODE.py <beta>") integrates the two ODEs and prints
// "[Zs] [Ys]" (GoldbeterModel_torus.cpp:254-261, util/GoldbeterModel/SolveGoldbeterODE.py).  The fixed point itself needs no
// integration: the sum of the two equations gives Zs = (v0 + v1 beta)/k, and Ys is the root of v2(Zs) - v3(Zs, Y) - kf Y = 0
// (monotone in Y: bisection).  The script protocol is kept as an option (ini key System.steadyStateCommand) for runs that must
// start from the script's own values, which carry its integration error (~1e-6).
#pragma once
#include <cstdio>
#include <string>

namespace crd {

inline void goldbeter_steady_state(double beta, double &Zs, double &Ys) {
  const double v0 = 1.0, k = 10.0, kf = 1.0, v1 = 7.3, VM2 = 65.0, VM3 = 500.0, K2 = 1.0, KR = 2.0, KA = 0.9;
  Zs = (v0 + v1 * beta) / k;
  const double z2 = Zs * Zs, z4 = z2 * z2;
  const double v2 = VM2 * z2 / (K2 * K2 + z2);
  auto g = [&](double Y) { return v2 - VM3 * Y * Y * z4 / ((KR * KR + Y * Y) * (KA * KA * KA * KA + z4)) - kf * Y; };
  double lo = 0.0, hi = 1.0;
  while (g(hi) > 0.0 && hi < 1e6) hi *= 2.0;
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (g(mid) > 0.0) lo = mid; else hi = mid;
  }
  Ys = 0.5 * (lo + hi);
}

// the reference's protocol: run `<command> <beta as written in the ini file>` and read "[Zs] [Ys]" from its output;
// false when the command cannot be run or prints something else
inline bool goldbeter_steady_state_from_command(const std::string &command, const std::string &beta_text, double &Zs, double &Ys) {
  const std::string line = command + " " + beta_text;
  FILE *in = popen(line.c_str(), "r");
  if (!in) return false;
  double z = 0, y = 0;
  const int got = fscanf(in, " [%lf] [%lf]", &z, &y);
  const int status = pclose(in);
  if (got != 2 || status != 0) return false;
  Zs = z; Ys = y;
  return true;
}

}  // namespace crd
