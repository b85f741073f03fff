// Host check of the float-vs-double threshold helpers of csrc/common.cuh (TEST INFRASTRUCTURE).
// torchvision compares a float IoU with the DOUBLE threshold (`ovr > thr`); the kernels compare floats only, against
// f = float_at_or_below(thr), which is valid iff  ((double)x > thr) <=> (x > f)  for EVERY float x -- checked here for the
// floats around f, for random thresholds, thresholds that are exactly floats, and thresholds one double ulp off a float.
// Same for float_at_or_above and `>=` (the rotated API's mode).  Prints the number of violations.
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <random>
#include <cuda_runtime.h>
#include "../../mydetection_b200/csrc/common.cuh"

static long check(double t) {
    long bad = 0;
    const float fb = mydet::float_at_or_below(t), fa = mydet::float_at_or_above(t);
    float xs[8] = {std::nextafterf(std::nextafterf(fb, -INFINITY), -INFINITY), std::nextafterf(fb, -INFINITY), fb,
                   std::nextafterf(fb, INFINITY), std::nextafterf(fa, -INFINITY), fa, std::nextafterf(fa, INFINITY),
                   std::nextafterf(std::nextafterf(fa, INFINITY), INFINITY)};
    for (float x : xs) {
        if (((double)x > t) != (x > fb)) ++bad;
        if (((double)x >= t) != (x >= fa)) ++bad;
    }
    return bad;
}

int main() {
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    long bad = 0, n = 0;
    for (int i = 0; i < 2000000; ++i) {
        const double t = u(rng);
        const float f = (float)t;
        const double cases[5] = {t, (double)f, std::nextafter((double)f, 2.0), std::nextafter((double)f, -1.0), t * 1e-3};
        for (double c : cases) { bad += check(c); ++n; }
    }
    for (double c : {0.0, 1.0, 0.45, 0.5, 0.3, 0.7, 1.0 / 3.0, 1.0 / 9.0, 1e-30, 0.9999999999}) { bad += check(c); ++n; }
    printf("%ld thresholds, %ld violations\n", n, bad);
    return bad ? 1 : 0;
}
