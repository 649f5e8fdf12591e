// TEST INFRASTRUCTURE.  One label map per input line; for each, the row assignment through the in-place
// view (FlatLabels) and through the DOM parser (Json) on two independent Assigners.  Prints per line:
//   <flat accepted 0|1> <dom result> <dom pod> <dom slot> [<flat result> <flat pod> <flat slot>]
// where result is S(kipped) / H (shadowed) / P(laced) / E (DOM parse error).
#include <cstdio>
#include <fstream>
#include <string>

#include "ingest_internal.hpp"

using namespace gph;
using namespace gph::detail;

static char code(Assigner::Result r) { return r == Assigner::Placed ? 'P' : r == Assigner::Shadowed ? 'H' : 'S'; }

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const bool is_power = argv[2][0] == '1';
  std::ifstream f(argv[1]);
  std::string line;
  Window wd, wf;
  Assigner ad(wd), af(wf);
  FlatLabels flat;
  while (std::getline(f, line)) {
    uint32_t p = 0, s = 0;
    const bool ok = flat.parse(line.data(), line.data() + line.size());
    Json j;
    bool dom_ok = true;
    try {
      j = Json::parse(line);
    } catch (const std::exception&) {
      dom_ok = false;
    }
    if (!dom_ok || !j.is_object()) {
      printf("%d E 0 0\n", (int)ok);
      continue;
    }
    const Assigner::Result rd = ad.assign(j, is_power, false, &p, &s);
    printf("%d %c %u %u", (int)ok, code(rd), rd == Assigner::Placed ? p : 0, rd == Assigner::Placed ? s : 0);
    if (ok) {
      p = s = 0;
      const Assigner::Result rf = af.assign(flat, is_power, false, &p, &s);
      printf(" %c %u %u", code(rf), rf == Assigner::Placed ? p : 0, rf == Assigner::Placed ? s : 0);
    } else {  // keep the second assigner in step: it sees the map through the DOM as the product does
      af.assign(j, is_power, false, &p, &s);
    }
    printf("\n");
  }
  // the two windows must have ended up with the same pods and slots
  bool same = wd.pods.size() == wf.pods.size() && wd.stats.series_skipped == wf.stats.series_skipped &&
              wd.stats.duplicates_merged == wf.stats.duplicates_merged;
  for (size_t i = 0; same && i < wd.pods.size(); ++i) {
    same = wd.pods[i].name == wf.pods[i].name && wd.pods[i].ns == wf.pods[i].ns &&
           wd.pods[i].slots.size() == wf.pods[i].slots.size() && wd.pods[i].power_slots == wf.pods[i].power_slots;
    for (size_t k = 0; same && k < wd.pods[i].slots.size(); ++k) {
      const GpuSlot &a = wd.pods[i].slots[k], &b = wf.pods[i].slots[k];
      same = a.hostname == b.hostname && a.container == b.container && a.gpu == b.gpu && a.model == b.model &&
             a.node_type == b.node_type;
    }
  }
  printf("%s\n", same ? "WINDOWS_EQUAL" : "WINDOWS_DIFFER");
  return same ? 0 : 1;
}
