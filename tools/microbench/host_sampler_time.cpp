// Host-side timing of tlod_anchor_subsample_host on this machine's CPU (no GPU work):
//   g++ -O3 -std=c++17 -I include -x c++ -c <pkg>/csrc/host_sampling.cu -o /tmp/hs.o
//   g++ -O3 -I include tools/microbench/host_sampler_time.cpp /tmp/hs.o -o /tmp/hs_time && /tmp/hs_time
#include <chrono>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "tlod_b200.h"

static double run(std::vector<float>& lab0, int B, int n, bool ahead_on) {
  std::vector<float> lab(lab0.size());
  std::mt19937 g(7);
  std::vector<unsigned> key(625);
  for (auto& k : key) k = g();
  int pos = 17, nex = 0;
  const int blocks = (int)(1.4 * B * n / 624) + 1;
  std::vector<unsigned> ahead((size_t)(blocks + 1) * 624);
  double best = 1e9;
  for (int it = 0; it < 300; ++it) {
    memcpy(lab.data(), lab0.data(), lab.size() * sizeof(float));
    if (ahead_on) tlod_mt_pregen(key.data(), ahead.data(), blocks);
    auto t0 = std::chrono::steady_clock::now();
    tlod_anchor_subsample_host_ahead(lab.data(), B, n, 128, 256, key.data(), &pos, ahead_on ? ahead.data() : nullptr,
                                     ahead_on ? blocks : 0, &nex);
    auto t1 = std::chrono::steady_clock::now();
    const double us = std::chrono::duration<double, std::micro>(t1 - t0).count();
    best = us < best ? us : best;
  }
  return best;
}

int main() {
  const int B = 2, n = 17434;
  std::vector<float> lab(B * n);
  std::mt19937 g(1);
  for (auto& v : lab) { unsigned r = g() % 1000; v = r < 30 ? -1.f : (r < 995 ? 0.f : 1.f); }
  printf("2 x %d labels, 96.5%% background: %.1f us; with key blocks generated ahead: %.1f us\n", n, run(lab, B, n, false),
         run(lab, B, n, true));
  std::vector<float> none(B * n, -1.f);
  printf("scan only (no label to subsample): %.1f us\n", run(none, B, n, false));
  std::vector<float> few(B * n, -1.f);
  for (int i = 0; i < B * n; i += 40) few[i] = 0.f;
  printf("2.5%% background (permutation of ~436 per image): %.1f us\n", run(few, B, n, false));
}
