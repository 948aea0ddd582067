// Mirror of the reference's only driver (CentoidMPCTest.cpp:11-116) against the C++ shim:
// same constructor arguments, SetupMPC(), one UpdateMPC() on the same fixture values, and
// -- unlike the reference, which asserts nothing -- prints the solution as one JSON line so
// that tests/test_gpu_shim.py can compare it with the CPU oracle.  Also exercises the
// "mpc table invalid" exception (reference CentroidalMPC.cpp:328-330).
#include <cstdio>
#include <vector>

#include "../../cheeta-mpc_b200/include/CentroidalMPC.h"

int main() {
  const double mass = 8, time_step = 0.01;
  const int num_legs = 4, horizon = 6;
  std::vector<double> mu = {0.8, 0.8, 0.8, 0.8};
  std::vector<double> weights = {1, 1, 100, 0.5, 0.5, 0, 2, 2, 8};
  for (int i = 0; i < 4; ++i) { const double w9[9] = {0.2, 0.2, 0.2, 0.3, 0.3, 0.3, 0.1, 0.1, 0.1}; weights.insert(weights.end(), w9, w9 + 9); }
  CentroidalMPC mpc(mass, num_legs, horizon, time_step, weights, mu);
  mpc.SetupMPC();
  std::vector<double> state = {0, 0, 0.15, 0.1, 0, 0, 0, 0, 0.1, 0.35, 0.052, 0, 0.35, -0.054, 0, -0.37, -0.053, 0, -0.36, 0.054, 0};
  std::vector<double> des_state(9 * (horizon + 1), 0.0);
  const double given[54] = {0.31, 0, 0.16, 0.32, 0, 0.168, 0.33, 0, 0.172, 0.33, 0, 0.18, 0.34, 0, 0.19, 0.348, 0, 0.2,
                            0.1, 0, 0, 0.09, 0, 0, 0.08, 0, 0, 0.06, 0, 0, 0.04, 0, 0, 0, 0, 0,
                            0, 0, 0.12, 0, 0, 0.14, 0, 0, 0.16, 0, 0, 0.18, 0, 0, 0.2, 0, 0, 0.22};
  for (int i = 0; i < 54; ++i) des_state[i] = given[i];  // the driver's comma initialiser fills 54 of 63 slots
  const double table[6][4] = {{1, 0, 1, 0}, {1, 0, 1, 0}, {1, 0, 1, 0}, {0, 1, 0, 1}, {0, 1, 0, 1}, {0, 1, 0, 1}};
  const double feet[4][7][3] = {
      {{0.35, 0.052, 0}, {0.35, 0.052, 0}, {0.35, 0.052, 0}, {0.35, 0.052, 0}, {0.38, 0.052, 0}, {0.39, 0.052, 0}, {0.42, 0.052, 0}},
      {{0.35, -0.054, 0}, {0.37, -0.052, 0}, {0.39, -0.052, 0}, {0.43, -0.052, 0}, {0.43, -0.052, 0}, {0.43, -0.052, 0}, {0.43, -0.052, 0}},
      {{-0.37, -0.052, 0}, {-0.37, -0.052, 0}, {-0.37, -0.052, 0}, {-0.36, -0.052, 0}, {-0.34, -0.052, 0}, {-0.30, -0.052, 0}, {-0.28, -0.052, 0}},
      {{-0.36, 0.053, 0}, {-0.34, 0.053, 0}, {-0.32, 0.053, 0}, {-0.31, 0.053, 0}, {-0.31, 0.052, 0}, {-0.31, 0.052, 0}, {-0.31, 0.052, 0}}};
  std::vector<double> des_input(num_legs * (4 * horizon + 3), 0.0);
  for (int i = 0; i < num_legs; ++i) {
    const int o = i * (4 * horizon + 3);
    for (int j = 0; j < horizon; ++j) des_input[o + j] = table[j][i];
    for (int k = 0; k <= horizon; ++k)
      for (int a = 0; a < 3; ++a) des_input[o + horizon + 3 * k + a] = feet[i][k][a];
  }
  std::vector<double> f = mpc.UpdateMPC(state, des_state, des_input);
  std::printf("{\"forces\": [");
  for (size_t i = 0; i < f.size(); ++i) std::printf("%s%.17g", i ? ", " : "", f[i]);
  std::printf("], ");
  bool threw = false;
  for (int i = 0; i < num_legs; ++i) des_input[i * (4 * horizon + 3) + 2] = 0.0;  // step 2: no stance leg
  try { mpc.UpdateMPC(state, des_state, des_input); } catch (const std::runtime_error& e) { threw = std::string(e.what()) == "mpc table invalid"; }
  std::printf("\"invalid_table_throws\": %s}\n", threw ? "true" : "false");
  std::printf("finished test\n");
  return 0;
}
