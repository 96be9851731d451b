// Host-only check of the forward DWT work queue (dwt_fast.cu build_queue): every item exactly once, every item after
// ALL strips of the chunks it reads (ticket order is a topological order -> the persistent kernel cannot deadlock), and
// the dependency ranges cover every row the item loads.  Compiled and run by tests/test_dwt_queue_order.py.
#include "dwt_fast.cu"
#include <cstdio>
#include <set>
using namespace pcs;
int check(int Wa, int Ha, int nl, int frames, bool lossy, int rp1, int lag) {
  QueueGeo g[kMaxQLevels];
  int maxc = queue_geo(Wa, Ha, nl, rp1, g);
  auto items = build_queue(g, nl, frames, lossy, lag);
  long long expect = 0;
  for (int l = 0; l < nl; ++l) expect += (long long)g[l].nstrips * g[l].nchunks * frames;
  if ((long long)items.size() != expect) { printf("count mismatch %zu vs %lld\n", items.size(), expect); return 1; }
  std::map<std::tuple<int,int,int>, std::pair<int,long long>> done;  // (f,l,c) -> (strips placed, last ticket)
  std::set<std::tuple<int,int,int,int>> seen;
  const int warm = lossy ? 2 : 1, delay = lossy ? 1 : 0;
  long long minlag = 1 << 30; long long sumlag = 0, ndep = 0;
  for (size_t t = 0; t < items.size(); ++t) {
    int l = items[t].x & 0xff, f = items[t].x >> 8, s = items[t].y, c = items[t].z, lo = items[t].w & 0xffff, hi = items[t].w >> 16;
    if (!seen.insert({f, l, c, s}).second) { printf("dup\n"); return 1; }
    if (l >= nl || s >= g[l].nstrips || c >= g[l].nchunks || f >= frames) { printf("range\n"); return 1; }
    if (l) {
      // rows needed
      int n0 = c * g[l].rp, n1 = std::min(n0 + g[l].rp, g[l].hh), Hc = 2 * g[l].hh;
      for (int y = 2 * (n0 - warm); y <= 2 * (n1 - 1 + delay) + 2; ++y) {
        int ys = y < 0 ? -y : (y >= Hc ? 2 * (Hc - 1) - y : y);
        int ch = ys / g[l - 1].rp;
        if (ch < lo || ch > hi) { printf("row %d (chunk %d) outside deps [%d,%d] level %d chunk %d\n", ys, ch, lo, hi, l, c); return 1; }
      }
      for (int d = lo; d <= hi; ++d) {
        auto it = done.find({f, l - 1, d});
        if (it == done.end() || it->second.first != g[l - 1].nstrips) { printf("dep not complete: t=%zu l=%d c=%d dep=%d\n", t, l, c, d); return 1; }
        long long lg = (long long)t - it->second.second; minlag = std::min(minlag, lg); 
      }
      auto it = done.find({f, l - 1, hi}); sumlag += (long long)t - it->second.second; ++ndep;
    }
    auto& e = done[{f, l, c}]; e.first++; e.second = t;
  }
  printf("%dx%d nl=%d frames=%d lossy=%d rp1=%d lag=%d: %zu items, maxchunks %d, min lag %lld, mean lag %.0f OK\n", Wa, Ha, nl, frames, lossy, rp1, lag, items.size(), maxc, minlag, ndep ? (double)sumlag / ndep : 0.0);
  return 0;
}
namespace pcs { unsigned long long g_launches = 0; }
int main() {
  int bad = 0;
  bad |= check(4096, 4096, 5, 16, true, 64, 2960);
  bad |= check(4096, 4096, 5, 2, true, 64, 2960);
  bad |= check(4096, 4096, 5, 1, false, 64, 2960);
  bad |= check(7680, 4352, 6, 1, false, 64, 2960);
  bad |= check(3840, 2176, 5, 8, true, 64, 2960);
  bad |= check(512, 512, 3, 1, true, 64, 2960);
  bad |= check(512, 512, 3, 3, false, 8, 0);
  bad |= check(256, 192, 2, 1, true, 64, 100);
  bad |= check(320, 128, 2, 5, false, 16, 7);
  bad |= check(1024, 64, 2, 4, true, 64, 50);
  bad |= check(32768, 4096, 5, 1, false, 64, 2960);
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}
