// Host build of gecco_b200/csrc/gcrf_exp.cuh: prints exp_cr(x) and libm's exp(x) as hex doubles for every x on stdin
// (one hex or decimal double per line).  tests/test_exp_dd.py compares both with 60-digit decimals.
#include <cstdio>
#include <cstdlib>

#include "../gecco_b200/csrc/gcrf_exp.cuh"

int main() {
    char line[256];
    while (fgets(line, sizeof line, stdin)) {
        const double x = strtod(line, nullptr);
        printf("%a %a\n", gcrf::expdd::exp_cr(x), exp(x));
    }
    return 0;
}
