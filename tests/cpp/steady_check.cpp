// steady_check.cpp — test program for crdmodel_b200/host/crd_steady.hpp.
// usage: steady_check closed <beta>            prints Zs Ys and the residuals of the two ODE right-hand sides
//        steady_check command <cmd> <beta>     prints ok Zs Ys | failed
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "crd_steady.hpp"

int main(int argc, char **argv) {
  if (argc == 3 && !std::strcmp(argv[1], "closed")) {
    const double beta = std::atof(argv[2]);
    double Z, Y;
    crd::goldbeter_steady_state(beta, Z, Y);
    // GoldbeterModel_torus.cpp:694-695,715-716
    const double v2 = 65.0 * std::pow(Z, 2.0) / (std::pow(1.0, 2.0) + std::pow(Z, 2.0));
    const double v3 = 500.0 * std::pow(Y, 2.0) * std::pow(Z, 4.0) / ((std::pow(2.0, 2.0) + std::pow(Y, 2.0)) * (std::pow(0.9, 4.0) + std::pow(Z, 4.0)));
    std::printf("%.17g %.17g %.3e %.3e\n", Z, Y, 1.0 + 7.3 * beta - v2 + v3 + 1.0 * Y - 10.0 * Z, v2 - v3 - 1.0 * Y);
    return 0;
  }
  if (argc == 4 && !std::strcmp(argv[1], "command")) {
    double Z = -1, Y = -1;
    if (crd::goldbeter_steady_state_from_command(argv[2], argv[3], Z, Y)) std::printf("ok %.17g %.17g\n", Z, Y);
    else std::printf("failed\n");
    return 0;
  }
  return 2;
}
