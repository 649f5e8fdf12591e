// gpu-pruner — C++ host of the B200 idle-decision engine, with the reference controller's
// command-line surface (/root/reference/gpu-pruner/src/main.rs:46-134, 273-375).
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "cli.hpp"
#include "controller.hpp"
#include "kube.hpp"
#include "promql.hpp"

int main(int argc, char** argv) {
  std::vector<std::string> args(argv + 1, argv + argc);
  gph::ParseOutcome po = gph::parse_cli(args);
  if (!po.ok) {
    fputs(po.message.c_str(), po.exit_code == 0 ? stdout : stderr);
    return po.exit_code;
  }
  const gph::Cli& cli = po.cli;
  if (cli.print_query) {  // the text the reference logs as "Running w/ Query: ..." (main.rs:282)
    fputs(gph::render_query(cli).c_str(), stdout);
    return 0;
  }
  gph::Logger log(cli.log_format, stderr);
  log.info("Enabled resources: " + std::to_string((int)gph::get_enabled_resources(cli.enabled_resources)));
  log.info("Running w/ Query: " + gph::render_query(cli));
  const gph::Selectors sel = gph::render_selectors(cli);
  log.info("Engine selectors: " + sel.util + (sel.power.empty() ? "" : " ; " + sel.power));

  std::unique_ptr<gph::FixtureKubeApi> kube;
  if (cli.kube_fixture) kube = std::make_unique<gph::FixtureKubeApi>(*cli.kube_fixture);
  std::unique_ptr<gph::VerdictEngine> engine = gph::make_gpr_engine();   // libgpr.so; no CPU fallback
  // The response text is parsed on the GPU straight into HBM (ingest_device.hpp); responses that are not
  // in Prometheus' compact encoding fall back to the CPU text parser by themselves.  GPR_INGEST=cpu forces
  // the threaded CPU text parser (window uploaded by gpr_decide) for comparison.
  const char* ing = getenv("GPR_INGEST");
  std::unique_ptr<gph::TextIngestor> cpu_ingestor;
  gph::TextIngestor* ingestor = engine->text_ingestor();
  if (ing && std::string(ing) == "cpu") cpu_ingestor = gph::make_cpu_text_ingestor(), ingestor = cpu_ingestor.get();
  std::unique_ptr<gph::WindowSource> src = gph::make_window_source(cli.prometheus_url, ingestor, &log);
  gph::Controller ctl(cli, kube.get(), engine.get(), log, gph::system_clock());
  return ctl.run(*src);
}
