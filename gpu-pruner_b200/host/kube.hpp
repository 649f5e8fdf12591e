// kube.hpp — the controller's Kubernetes domain logic as pure functions over JSON objects.
//
// Mirrors /root/reference/gpu-pruner/src/lib.rs: ScaleKind (lib.rs:36-43) with its Eq/Hash by
// variant + UID (lib.rs:45-82), ResourceKind bitflags (lib.rs:96-105), get_enabled_resources
// (lib.rs:116-129), Meta (lib.rs:299-335), generate_scale_event (lib.rs:388-427),
// find_root_object (lib.rs:437-513) and the three scale-to-zero patches (lib.rs:517-576).
// The HTTP client itself is out of scope (no cluster, no network): every lookup goes through the
// KubeApi interface, implemented here over a directory of JSON fixtures, and every mutation is
// returned as a Request {method, path, content-type, body} that a transport would send.
#pragma once
#include <cstdint>
#include <functional>
#include <optional>
#include <string>
#include <vector>

#include "json.hpp"

namespace gph {

// ---- ResourceKind bitflags (lib.rs:96-105) ----------------------------------------------------
enum ResourceKind : uint8_t {
  RK_NONE = 0,
  RK_DEPLOYMENT = 0b00001,
  RK_REPLICA_SET = 0b00010,
  RK_STATEFUL_SET = 0b00100,
  RK_INFERENCE_SERVICE = 0b01000,
  RK_NOTEBOOK = 0b10000,
};
// 'd','r','s','i','n'; unknown characters are silently ignored (lib.rs:116-129)
uint8_t get_enabled_resources(const std::string& letters);

// ---- ScaleKind --------------------------------------------------------------------------------
enum class Kind { Deployment, ReplicaSet, StatefulSet, InferenceService, Notebook };

struct ScaleKind {
  Kind kind;
  Json object;  // the fetched resource

  // Meta (lib.rs:299-335)
  std::string name() const;
  std::optional<std::string> ns() const;
  std::string kind_name() const;      // "Deployment", "ReplicaSet", "StatefulSet", "Notebook", "InferenceService"
  std::string api_version() const;    // "apps/v1" x3, "v1" (Notebook), "v1beta1" (InferenceService)
  std::optional<std::string> uid() const;
  std::optional<std::string> resource_version() const;
  uint8_t resource_kind() const;      // From<ScaleKind> for ResourceKind (lib.rs:84-94)

  // Eq: same variant and, for Deployment/ReplicaSet/StatefulSet, identical objects; for
  // InferenceService/Notebook, identical UIDs (lib.rs:45-59).  Hash: variant + UID (lib.rs:62-82).
  bool operator==(const ScaleKind& o) const;
  size_t hash() const;
};
struct ScaleKindHash {
  size_t operator()(const ScaleKind& s) const { return s.hash(); }
};

// ---- API access ---------------------------------------------------------------------------------
struct Request {
  std::string method;        // POST / PATCH
  std::string path;          // e.g. /apis/apps/v1/namespaces/ns/deployments/name/scale
  std::string content_type;  // application/json | application/merge-patch+json
  Json body;
};

class KubeApi {
 public:
  virtual ~KubeApi() = default;
  // nullopt = 404; throws std::runtime_error for transport-level failures
  virtual std::optional<Json> get(Kind k, const std::string& ns, const std::string& name) = 0;
  virtual std::optional<Json> get_pod(const std::string& ns, const std::string& name) = 0;
};

// Directory layout: <dir>/<plural>/<namespace>/<name>.json with plural in
// {pods, replicasets, deployments, statefulsets, notebooks, inferenceservices}
class FixtureKubeApi : public KubeApi {
 public:
  explicit FixtureKubeApi(std::string dir) : dir_(std::move(dir)) {}
  std::optional<Json> get(Kind k, const std::string& ns, const std::string& name) override;
  std::optional<Json> get_pod(const std::string& ns, const std::string& name) override;
  uint64_t calls = 0;

 private:
  std::optional<Json> load(const std::string& plural, const std::string& ns, const std::string& name);
  std::string dir_;
};

// ---- owner walk (lib.rs:437-513) -------------------------------------------------------------------
struct RootResult {
  std::optional<ScaleKind> root;
  std::string error;  // set when root is empty: "no scalable root object found for pod ..." or a lookup error
};
RootResult find_root_object(KubeApi& api, const Json& pod_metadata);

// ---- mutations (lib.rs:337-427, 517-576) -------------------------------------------------------------
struct Clock {
  std::function<int64_t()> now_ns;                 // wall clock, injectable for tests
  std::function<std::string()> uuid_simple;        // 32 hex digits
};
Clock system_clock();
std::string rfc3339(int64_t unix_ns);              // jiff Timestamp Display: 2024-01-02T03:04:05.123456789Z
int64_t parse_rfc3339(const std::string& s);       // -> unix ns; throws on garbage

Json generate_scale_event(const ScaleKind& sk, const Clock& clock, const std::string& pod_name_env);
// Event POST (if namespaced) followed by the scale request, as ScaleKind::scale issues them
std::vector<Request> scale_requests(const ScaleKind& sk, const Clock& clock,
                                    const std::string& pod_name_env);

const char* plural_of(Kind k);
std::string api_path(Kind k, const std::string& ns, const std::string& name);

}  // namespace gph
