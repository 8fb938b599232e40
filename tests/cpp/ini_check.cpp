#include "crd_ini.hpp"
#include <cstdio>
#include <cstring>
// Test helper: the lookup contract the reference gets from Boost.PropertyTree (src/FHNmodel_torus.cpp:158-174):
// [Section] headers, key = value, comment lines, missing key / bad conversion throw, the shipped ini files parse.
#define CHECK(x) do { if (!(x)) { std::fprintf(stderr, "ini_check: failed: %s (line %d)\n", #x, __LINE__); return 1; } } while (0)
int main(int argc, char **argv) {
  if (argc != 2) return 2;
  crd::Ini pt(argv[1]);
  CHECK(pt.get<double>("Parameters.diffusion") == 0.12);
  CHECK(pt.get<int>("Parameters.xMesh") == 400);
  CHECK(pt.get<double>("Parameters.tFinal") == 50.0);
  CHECK(pt.get<int>("System.varyBeta") == 1);
  CHECK(pt.get<int>("System.gpus", 3) == 3);                 // absent: fallback
  CHECK(pt.has("Parameters.beta") && !pt.has("Parameters.thetaMesh") && !pt.has("beta"));
  CHECK(pt.str("Parameters.comment_like") == "a # b ; c");   // only whole-line comments are comments
  CHECK(pt.get<int>("top") == 7);                            // a key before any section has no prefix
  bool threw = false;
  try { pt.get<int>("Parameters.nope"); } catch (const std::runtime_error &e) { threw = std::strstr(e.what(), "No such node") != nullptr; }
  CHECK(threw);
  threw = false;
  try { pt.get<int>("Parameters.word"); } catch (const std::runtime_error &e) { threw = std::strstr(e.what(), "conversion") != nullptr; }
  CHECK(threw);
  // the whole value must convert (Boost's translator): no silent 1e3 -> 1 or 0.4abc -> 0.4
  for (const char *k : {"Parameters.sci_int", "Parameters.trailing", "Parameters.two_numbers"}) {
    threw = false;
    try { (k[11] == 's') ? (void)pt.get<int>(k) : (void)pt.get<double>(k); } catch (const std::runtime_error &e) { threw = std::strstr(e.what(), "conversion") != nullptr; }
    CHECK(threw);
  }
  CHECK(pt.get<double>("Parameters.sci_int") == 1000.0);     // as a double it is fine
  CHECK(pt.get<double>("Parameters.padded") == 2.5);         // surrounding blanks are not part of the value
  threw = false;
  try { crd::Ini missing("/nonexistent/file.ini"); } catch (const std::runtime_error &) { threw = true; }
  CHECK(threw);
  return 0;
}
