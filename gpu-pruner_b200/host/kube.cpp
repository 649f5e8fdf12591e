#include "kube.hpp"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <random>
#include <sstream>
#include <sys/stat.h>

namespace gph {

uint8_t get_enabled_resources(const std::string& letters) {
  uint8_t k = RK_NONE;
  for (char c : letters) {
    switch (c) {
      case 'd': k |= RK_DEPLOYMENT; break;
      case 'r': k |= RK_REPLICA_SET; break;
      case 's': k |= RK_STATEFUL_SET; break;
      case 'i': k |= RK_INFERENCE_SERVICE; break;
      case 'n': k |= RK_NOTEBOOK; break;
      default: break;  // unknown characters are silently ignored
    }
  }
  return k;
}

// ---- Meta -------------------------------------------------------------------------------------------
static std::optional<std::string> opt_str(const Json& j) {
  if (j.is_string()) return j.as_string();
  return std::nullopt;
}
std::string ScaleKind::name() const { return object["metadata"]["name"].as_string(); }
std::optional<std::string> ScaleKind::ns() const { return opt_str(object["metadata"]["namespace"]); }
std::optional<std::string> ScaleKind::uid() const { return opt_str(object["metadata"]["uid"]); }
std::optional<std::string> ScaleKind::resource_version() const {
  return opt_str(object["metadata"]["resourceVersion"]);
}
std::string ScaleKind::kind_name() const {
  switch (kind) {
    case Kind::Deployment: return "Deployment";
    case Kind::ReplicaSet: return "ReplicaSet";
    case Kind::StatefulSet: return "StatefulSet";
    case Kind::Notebook: return "Notebook";
    default: return "InferenceService";
  }
}
std::string ScaleKind::api_version() const {  // lib.rs:308-316
  switch (kind) {
    case Kind::Notebook: return "v1";
    case Kind::InferenceService: return "v1beta1";
    default: return "apps/v1";
  }
}
uint8_t ScaleKind::resource_kind() const {
  switch (kind) {
    case Kind::Deployment: return RK_DEPLOYMENT;
    case Kind::ReplicaSet: return RK_REPLICA_SET;
    case Kind::StatefulSet: return RK_STATEFUL_SET;
    case Kind::InferenceService: return RK_INFERENCE_SERVICE;
    default: return RK_NOTEBOOK;
  }
}
bool ScaleKind::operator==(const ScaleKind& o) const {
  if (kind != o.kind) return false;
  if (kind == Kind::InferenceService || kind == Kind::Notebook) return uid() == o.uid();
  return object.dump() == o.object.dump();  // derived PartialEq on the whole resource
}
size_t ScaleKind::hash() const {
  const std::string u = uid().value_or(std::string("\x01none"));
  return std::hash<std::string>()(u) * 1000003u + (size_t)kind;
}

// ---- fixtures --------------------------------------------------------------------------------------------
const char* plural_of(Kind k) {
  switch (k) {
    case Kind::Deployment: return "deployments";
    case Kind::ReplicaSet: return "replicasets";
    case Kind::StatefulSet: return "statefulsets";
    case Kind::Notebook: return "notebooks";
    default: return "inferenceservices";
  }
}
std::string api_path(Kind k, const std::string& ns, const std::string& name) {
  std::string group;
  switch (k) {
    case Kind::Notebook: group = "/apis/kubeflow.org/v1"; break;                 // resources/src/notebook.rs:16-28
    case Kind::InferenceService: group = "/apis/serving.kserve.io/v1beta1"; break;  // inferenceservice.rs:16-31
    default: group = "/apis/apps/v1";
  }
  return group + "/namespaces/" + ns + "/" + plural_of(k) + "/" + name;
}

std::optional<Json> FixtureKubeApi::load(const std::string& plural, const std::string& ns,
                                         const std::string& name) {
  ++calls;
  const std::string path = dir_ + "/" + plural + "/" + ns + "/" + name + ".json";
  struct stat st;
  if (stat(path.c_str(), &st) != 0) return std::nullopt;
  return Json::parse_file(path);
}
std::optional<Json> FixtureKubeApi::get(Kind k, const std::string& ns, const std::string& name) {
  return load(plural_of(k), ns, name);
}
std::optional<Json> FixtureKubeApi::get_pod(const std::string& ns, const std::string& name) {
  return load("pods", ns, name);
}

// ---- owner walk -----------------------------------------------------------------------------------------------
RootResult find_root_object(KubeApi& api, const Json& meta) {
  RootResult r;
  const std::string pod_name = meta["name"].as_string();
  const std::string ns = meta["namespace"].as_string();  // unwrap_or_default
  // KServe shortcut (lib.rs:448-456): the label names the InferenceService directly; a failed GET
  // propagates (`?`)
  const Json& ks = meta["labels"]["serving.kserve.io/inferenceservice"];
  if (ks.is_string()) {
    auto is = api.get(Kind::InferenceService, ns, ks.as_string());
    if (!is) {
      r.error = "inferenceservices \"" + ks.as_string() + "\" not found";
      return r;
    }
    r.root = ScaleKind{Kind::InferenceService, *is};
    return r;
  }
  for (const Json& orf : meta["ownerReferences"].items()) {
    const std::string kind = orf["kind"].as_string();
    const std::string name = orf["name"].as_string();
    if (kind == "ReplicaSet") {
      std::optional<Json> rs;
      try {
        rs = api.get(Kind::ReplicaSet, ns, name);  // errors swallowed: `if let Ok(rs)` (lib.rs:465)
      } catch (const std::exception&) {
        rs.reset();
      }
      if (!rs) continue;
      for (const Json& o2 : (*rs)["metadata"]["ownerReferences"].items()) {
        if (o2["kind"].as_string() == "Deployment") {
          auto dep = api.get(Kind::Deployment, ns, o2["name"].as_string());  // `?` propagates
          if (!dep) {
            r.error = "deployments \"" + o2["name"].as_string() + "\" not found";
            return r;
          }
          r.root = ScaleKind{Kind::Deployment, *dep};
          return r;
        }
      }
      r.root = ScaleKind{Kind::ReplicaSet, *rs};  // replica set with no Deployment owner
      return r;
    } else if (kind == "StatefulSet") {
      std::optional<Json> ss;
      try {
        ss = api.get(Kind::StatefulSet, ns, name);
      } catch (const std::exception&) {
        ss.reset();
      }
      if (!ss) continue;
      for (const Json& o2 : (*ss)["metadata"]["ownerReferences"].items()) {
        if (o2["kind"].as_string() == "Notebook") {
          auto nb = api.get(Kind::Notebook, ns, o2["name"].as_string());
          if (!nb) {
            r.error = "notebooks \"" + o2["name"].as_string() + "\" not found";
            return r;
          }
          r.root = ScaleKind{Kind::Notebook, *nb};
          return r;
        }
      }
      r.root = ScaleKind{Kind::StatefulSet, *ss};
      return r;
    }
    // other kinds: ignored (lib.rs:502-504)
  }
  r.error = "no scalable root object found for pod " +
            (meta["name"].is_string() ? "Some(\"" + pod_name + "\")" : std::string("None"));
  return r;
}

// ---- time -------------------------------------------------------------------------------------------------------
std::string rfc3339(int64_t ns) {
  int64_t secs = ns / 1000000000ll, frac = ns % 1000000000ll;
  if (frac < 0) frac += 1000000000ll, --secs;
  time_t t = (time_t)secs;
  struct tm tm;
  gmtime_r(&t, &tm);
  char buf[64];
  strftime(buf, sizeof buf, "%Y-%m-%dT%H:%M:%S", &tm);
  std::string s = buf;
  if (frac) {  // jiff prints the shortest fraction: trailing zeros trimmed
    char f[16];
    snprintf(f, sizeof f, "%09lld", (long long)frac);
    std::string fs = f;
    while (!fs.empty() && fs.back() == '0') fs.pop_back();
    s += "." + fs;
  }
  return s + "Z";
}

int64_t parse_rfc3339(const std::string& s) {
  int Y, M, D, h, m;
  double sec;
  char tz[8] = "";
  if (sscanf(s.c_str(), "%d-%d-%dT%d:%d:%lf%7s", &Y, &M, &D, &h, &m, &sec, tz) < 6)
    throw std::runtime_error("bad RFC 3339 timestamp: " + s);
  struct tm tm;
  memset(&tm, 0, sizeof tm);
  tm.tm_year = Y - 1900, tm.tm_mon = M - 1, tm.tm_mday = D, tm.tm_hour = h, tm.tm_min = m;
  int64_t base = (int64_t)timegm(&tm);
  int64_t off = 0;
  if (tz[0] == '+' || tz[0] == '-') {
    int oh = 0, om = 0;
    sscanf(tz + 1, "%d:%d", &oh, &om);
    off = (oh * 3600 + om * 60) * (tz[0] == '+' ? 1 : -1);
  }
  const int64_t whole = (int64_t)sec;
  const int64_t frac = (int64_t)((sec - (double)whole) * 1e9 + 0.5);
  return (base - off + whole) * 1000000000ll + frac;
}

Clock system_clock() {
  Clock c;
  c.now_ns = [] {
    return (int64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
               std::chrono::system_clock::now().time_since_epoch()).count();
  };
  c.uuid_simple = [] {
    static std::mt19937_64 rng{std::random_device{}()};
    uint64_t a = rng(), b = rng();
    a = (a & 0xffffffffffff0fffull) | 0x0000000000004000ull;  // version 4
    b = (b & 0x3fffffffffffffffull) | 0x8000000000000000ull;  // variant 1
    char buf[40];
    snprintf(buf, sizeof buf, "%016llx%016llx", (unsigned long long)a, (unsigned long long)b);
    return std::string(buf);
  };
  return c;
}

// ---- Event + scale requests ----------------------------------------------------------------------------------------
static Json opt_json(const std::optional<std::string>& s) { return s ? Json(*s) : Json(); }

Json generate_scale_event(const ScaleKind& sk, const Clock& clock, const std::string& pod_name_env) {
  const int64_t now = clock.now_ns();
  const std::string ts = rfc3339(now / 1000000000ll * 1000000000ll);  // Time: second resolution
  // MicroTime serialises with exactly six fractional digits
  char micro[64];
  {
    time_t t = (time_t)(now / 1000000000ll);
    struct tm tm;
    gmtime_r(&t, &tm);
    char base[32];
    strftime(base, sizeof base, "%Y-%m-%dT%H:%M:%S", &tm);
    snprintf(micro, sizeof micro, "%s.%06lldZ", base, (long long)((now % 1000000000ll) / 1000));
  }
  const std::string ns = sk.ns().value_or("");
  Json ev = Json::object();
  ev.set("apiVersion", "v1");
  ev.set("kind", "Event");
  Json meta = Json::object();
  meta.set("name", "gpuscaler-" + clock.uuid_simple());
  if (sk.ns()) meta.set("namespace", *sk.ns());
  ev.set("metadata", meta);
  ev.set("action", "scale_down");
  ev.set("reason", "Pod " + ns + "::" + sk.name() + " was not using GPU");
  ev.set("type", "Normal");
  ev.set("reportingComponent", "gpu-pruner");
  ev.set("reportingInstance", pod_name_env.empty() ? std::string("gpu_pruner") : pod_name_env);
  ev.set("firstTimestamp", ts);
  ev.set("lastTimestamp", ts);
  ev.set("eventTime", std::string(micro));
  Json inv = Json::object();
  inv.set("apiVersion", sk.api_version());
  inv.set("kind", sk.kind_name());
  inv.set("name", sk.name());
  if (sk.ns()) inv.set("namespace", *sk.ns());
  if (sk.resource_version()) inv.set("resourceVersion", *sk.resource_version());
  if (sk.uid()) inv.set("uid", *sk.uid());
  ev.set("involvedObject", inv);
  (void)opt_json;
  return ev;
}

std::vector<Request> scale_requests(const ScaleKind& sk, const Clock& clock,
                                    const std::string& pod_name_env) {
  std::vector<Request> out;
  const std::string ns = sk.ns().value_or("");
  if (sk.ns()) {  // Event first; its failure is logged, not fatal (lib.rs:340-349)
    out.push_back(Request{"POST", "/api/v1/namespaces/" + ns + "/events", "application/json",
                          generate_scale_event(sk, clock, pod_name_env)});
  }
  Request rq;
  rq.method = "PATCH";
  rq.content_type = "application/merge-patch+json";
  const std::string base = api_path(sk.kind, ns, sk.name());
  switch (sk.kind) {
    case Kind::Deployment:
    case Kind::ReplicaSet:
    case Kind::StatefulSet: {  // lib.rs:517-525: /scale subresource, {"spec":{"replicas":0}}
      rq.path = base + "/scale";
      Json spec = Json::object();
      spec.set("replicas", 0);
      rq.body = Json::object();
      rq.body.set("spec", spec);
      break;
    }
    case Kind::Notebook: {  // lib.rs:529-549: stop annotation with the current time
      rq.path = base;
      Json ann = Json::object();
      ann.set("kubeflow-resource-stopped", rfc3339(clock.now_ns()));
      Json meta = Json::object();
      meta.set("annotations", ann);
      rq.body = Json::object();
      rq.body.set("metadata", meta);
      break;
    }
    case Kind::InferenceService: {  // lib.rs:553-576: spec.predictor.minReplicas = 0
      rq.path = base;
      Json pred = Json::object();
      pred.set("minReplicas", 0);
      Json spec = Json::object();
      spec.set("predictor", pred);
      rq.body = Json::object();
      rq.body.set("spec", spec);
      break;
    }
  }
  out.push_back(rq);
  return out;
}

}  // namespace gph
