#include "promql.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace gph {

LabelNames label_names(bool honor) {
  return honor ? LabelNames{"pod", "namespace", "container"}
               : LabelNames{"exported_pod", "exported_namespace", "exported_container"};
}

// shortest decimal that round-trips, with a ".0" appended to integral values — what both
// minijinja and Python print for an f64 (`--power-threshold 150` renders as `>= 150.0`)
std::string format_float(double v) {
  if (std::isnan(v)) return "nan";
  if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
  if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
  // shortest digit string that round-trips (what repr / minijinja print)
  char buf[48];
  int prec = 1;
  for (; prec <= 17; ++prec) {
    snprintf(buf, sizeof buf, "%.*e", prec - 1, v);
    if (strtod(buf, nullptr) == v) break;
  }
  // buf = [-]d.ddddde[+-]XX
  std::string m = buf;
  const size_t epos = m.find('e');
  const int exp10 = atoi(m.c_str() + epos + 1);
  std::string digits;
  bool neg = false;
  for (size_t k = 0; k < epos; ++k) {
    if (m[k] == '-') neg = true;
    else if (m[k] != '.') digits += m[k];
  }
  std::string out = neg ? "-" : "";
  if (exp10 >= -4 && exp10 < 16) {  // positional notation, always with a fractional part
    if (exp10 >= 0) {
      std::string ip = digits.substr(0, std::min<size_t>(digits.size(), (size_t)exp10 + 1));
      while ((int)ip.size() < exp10 + 1) ip += '0';
      std::string fp = digits.size() > (size_t)exp10 + 1 ? digits.substr((size_t)exp10 + 1) : "0";
      out += ip + "." + fp;
    } else {
      out += "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
    }
  } else {  // scientific: d[.ddd]e[+-]XX with at least two exponent digits
    out += digits.substr(0, 1);
    if (digits.size() > 1) out += "." + digits.substr(1);
    char e[16];
    snprintf(e, sizeof e, "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
    out += e;
  }
  return out;
}

namespace {

// the selector body shared by every metric: `pl != ""[, nl =~ "ns"][, modelName =~ "m"]`
std::string matchers(const Cli& a, bool with_model) {
  const LabelNames l = label_names(a.honor_labels);
  std::string s = std::string(l.pod) + " != \"\"";
  if (a.ns && !a.ns->empty()) s += std::string(", ") + l.ns + " =~ \"" + *a.ns + "\"";
  if (with_model && a.model_name && !a.model_name->empty())
    s += ", modelName =~ \"" + *a.model_name + "\"";
  return s;
}

std::string by_labels(const Cli& a) {
  const LabelNames l = label_names(a.honor_labels);
  return std::string("Hostname, ") + l.container + ", " + l.pod + ", " + l.ns + ", gpu, modelName";
}

std::string idle_gpus(const Cli& a) {  // query.promql.j2:8-22
  const std::string d = std::to_string(a.duration);
  const std::string m = matchers(a, true);
  return "sum by (" + by_labels(a) + ") (\n"
         "    max_over_time(DCGM_FI_PROF_GR_ENGINE_ACTIVE{\n"
         "      " + m + "\n"
         "    }[" + d + "m])\n"
         "    or\n"
         "    max_over_time(DCGM_FI_DEV_GPU_UTIL{\n"
         "      " + m + "\n"
         "    }[" + d + "m]) / 100\n"
         ")";
}

bool power_truthy(const Cli& a) {  // Jinja `{% if args.power_threshold %}`: None and 0.0 are falsy
  return a.power_threshold && *a.power_threshold != 0.0 && !std::isnan(*a.power_threshold);
}

}  // namespace

std::string render_query(const Cli& a) {
  const LabelNames l = label_names(a.honor_labels);
  const std::string ig = idle_gpus(a);
  std::string q =
      "(\n"
      "  " + ig + " * on (Hostname) group_left(node_type) (\n"
      "    label_replace(\n"
      "      label_replace(node_dmi_info,\n"
      "        \"Hostname\", \"$1\", \"instance\", \"(.+)\"\n"
      "      ),\n"
      "      \"node_type\", \"$1\", \"product_name\", \"(.+)\"\n"
      "    )\n"
      "  )\n"
      "  or on (" + by_labels(a) + ")\n"
      "  " + ig + "\n"
      ")\n"
      "== 0";
  if (power_truthy(a)) {
    q += "\nunless on (" + std::string(l.pod) + ", " + l.ns + ")\n"
         "(\n"
         "  max_over_time(DCGM_FI_DEV_POWER_USAGE{\n"
         "    " + matchers(a, false) + "\n"
         "  }[" + std::to_string(a.duration) + "m]) >= " + format_float(*a.power_threshold) + "\n"
         ")";
  }
  return q;
}

static Selectors selectors_over(const Cli& a, const std::string& d) {
  Selectors s;
  s.prof = "DCGM_FI_PROF_GR_ENGINE_ACTIVE{" + matchers(a, true) + "}" + d;
  s.util = "DCGM_FI_DEV_GPU_UTIL{" + matchers(a, true) + "}" + d;
  if (power_truthy(a)) s.power = "DCGM_FI_DEV_POWER_USAGE{" + matchers(a, false) + "}" + d;
  s.dmi = "node_dmi_info";
  return s;
}

Selectors render_selectors(const Cli& a) { return selectors_over(a, "[" + std::to_string(a.duration) + "m]"); }

Selectors render_selectors(const Cli& a, int64_t seconds) {
  return selectors_over(a, "[" + std::to_string(seconds) + "s]");
}

}  // namespace gph
