// Fuzz driver for the two ingest paths (built with -fsanitize=address,undefined by
// tests/test_ingest_fuzz.py).  For every input file: the text path and the DOM path must either both
// produce the same tensor or both reject the input with an exception; neither may crash, hang or read
// out of bounds.  Prints one line per file: OK / REJECT / MISMATCH.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>

#include "ingest.hpp"

using namespace gph;

static bool same(const std::vector<float>& a, const std::vector<float>& b) {
  return a.size() == b.size() && (a.empty() || memcmp(a.data(), b.data(), a.size() * sizeof(float)) == 0);
}

int main(int argc, char** argv) {
  int bad = 0;
  for (int i = 1; i < argc; ++i) {
    std::ifstream f(argv[i], std::ios::binary);
    std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    IngestOptions o;
    o.duration_min = 1;
    o.step = 1;
    o.t_end = 1700000000;
    bool ok_text = true, ok_dom = true;
    Window wt, wd;
    try {
      wt = ingest_matrix_text(s, nullptr, nullptr, o, 2);
    } catch (const std::exception&) {
      ok_text = false;
    }
    try {
      wd = ingest_matrix(Json::parse(s, 2), nullptr, nullptr, o);
    } catch (const std::exception&) {
      ok_dom = false;
    }
    const char* verdict = "OK";
    if (ok_text && ok_dom) {
      if (!(wt.P == wd.P && wt.G == wd.G && wt.T == wd.T && same(wt.util, wd.util))) verdict = "MISMATCH", ++bad;
    } else if (!ok_text && !ok_dom) {
      verdict = "REJECT";
    } else {
      // one path is stricter than the other on malformed input: acceptable only when the input is not
      // valid JSON for the DOM parser (the text path does not validate what it can skip over)
      verdict = ok_dom ? "MISMATCH" : "LENIENT";
      if (ok_dom) ++bad;
    }
    printf("%s %s\n", verdict, argv[i]);
  }
  return bad ? 1 : 0;
}
