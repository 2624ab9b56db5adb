// test_tables.cc — the time-stepping parameter tables of dune-pdelab_b200/host/onestep.hh printed as JSON lines; compared
// with python/pdelab_b200/onestep.py (and through it with the consistency conditions) by tests/test_cpp_partition.py.
#include <cstdio>

#include "../../dune-pdelab_b200/host/onestep.hh"

namespace PDELab = Dune::PDELab::B200;

static void dump(const char* cls, const PDELab::TimeSteppingParameterInterface<double>& m) {
  const int s = (int)m.s();
  std::printf("{\"class\": \"%s\", \"name\": \"%s\", \"implicit\": %s, \"s\": %d, \"a\": [", cls, m.name().c_str(),
              m.implicit() ? "true" : "false", s);
  for (int r = 1; r <= s; r++) {
    std::printf("%s[", r > 1 ? ", " : "");
    for (int i = 0; i <= r; i++) std::printf("%s%.17g", i ? ", " : "", m.a(r, i));
    std::printf("]");
  }
  std::printf("], \"b\": [");
  for (int r = 1; r <= s; r++) {
    std::printf("%s[", r > 1 ? ", " : "");
    for (int i = 0; i <= r; i++) std::printf("%s%.17g", i ? ", " : "", m.b(r, i));
    std::printf("]");
  }
  std::printf("], \"d\": [");
  for (int i = 0; i <= s; i++) std::printf("%s%.17g", i ? ", " : "", m.d(i));
  std::printf("]}\n");
}

int main() {
  dump("OneStepThetaParameter(0.5)", PDELab::OneStepThetaParameter<double>(0.5));
  dump("ExplicitEulerParameter", PDELab::ExplicitEulerParameter<double>());
  dump("ImplicitEulerParameter", PDELab::ImplicitEulerParameter<double>());
  dump("HeunParameter", PDELab::HeunParameter<double>());
  dump("Shu3Parameter", PDELab::Shu3Parameter<double>());
  dump("RK4Parameter", PDELab::RK4Parameter<double>());
  dump("Alexander2Parameter", PDELab::Alexander2Parameter<double>());
  dump("FractionalStepParameter", PDELab::FractionalStepParameter<double>());
  dump("Alexander3Parameter", PDELab::Alexander3Parameter<double>());
  return 0;
}
