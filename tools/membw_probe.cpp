// dev probe: multi-threaded memset bandwidth of the host (how fast could the host expand a sparse observation stream?)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
int main(int argc, char** argv) {
  const size_t total = (size_t)7400 << 20;
  char* buf = (char*)malloc(total);
  memset(buf, 1, total);
  for (int nt : {1, 4, 8, 16, 32}) {
    if (nt > (int)std::thread::hardware_concurrency()) break;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([=] { size_t per = total / nt; memset(buf + per * t, 0, per); });
    for (auto& x : th) x.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("threads %d: %.1f GB/s (%.1f ms for 7.4 GB)\n", nt, total / s / 1e9, s * 1e3);
  }
  printf("hw threads %u\n", std::thread::hardware_concurrency());
}
