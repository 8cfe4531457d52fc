// Row f1 measurement (round 2): in-process timing of MatchSuper4PCS::ComputeTransformation for several S4PCS_LANES
// values (and, through the caller's S4PCS_DEVICES, device-context counts), CUDA start-up excluded (one untimed warm-up run), median of R repetitions.  Uses only the public headers.
//   lanes_bench P.obj Q.obj overlap delta sample_size [reps=5] [lanes="1 2 4 8"]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <string>
#include <vector>

#include "super4pcs/algorithms/super4pcs.h"
#include "super4pcs/io/io.h"

using namespace GlobalRegistration;

static double run_once(const std::vector<Point3D>& P, const std::vector<Point3D>& Q0, const Match4PCSOptions& opt, float* score,
                       Match4PCSBase::MatrixType* T) {
  std::vector<Point3D> Q = Q0;
  Utils::Logger logger(Utils::NoLog);
  MatchSuper4PCS matcher(opt, logger);  // reads S4PCS_LANES
  *T = Match4PCSBase::MatrixType::Identity();
  const auto t0 = std::chrono::steady_clock::now();
  *score = matcher.ComputeTransformation(P, &Q, *T);
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char** argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s P.obj Q.obj overlap delta sample_size [reps] [\"lanes list\"]\n", argv[0]);
    return 2;
  }
  std::vector<Point3D> P, Q;
  {
    IOManager io;
    std::vector<Eigen::Matrix2f> tex;
    std::vector<Point3D::VectorType> nrm;
    std::vector<tripple> tris;
    std::vector<std::string> mtls;
    if (!io.ReadObject(argv[1], P, tex, nrm, tris, mtls)) return 3;
    tex.clear(); nrm.clear(); tris.clear(); mtls.clear();
    if (!io.ReadObject(argv[2], Q, tex, nrm, tris, mtls)) return 3;
  }
  Match4PCSOptions opt;
  opt.configureOverlap(std::atof(argv[3]));
  opt.delta = std::atof(argv[4]);
  opt.sample_size = std::atoi(argv[5]);
  opt.max_time_seconds = 100000;
  const int reps = argc > 6 ? std::atoi(argv[6]) : 5;
  std::istringstream lanes(argc > 7 ? argv[7] : "1 2 4 8");
  float score = 0, score0 = 0;
  Match4PCSBase::MatrixType T, T0;
  setenv("S4PCS_LANES", "1", 1);
  run_once(P, Q, opt, &score0, &T0);  // warm-up: CUDA context, module load, allocations
  for (int L; lanes >> L;) {
    setenv("S4PCS_LANES", std::to_string(L).c_str(), 1);
    std::vector<double> ms;
    bool same = true;
    for (int r = 0; r < reps; ++r) {
      ms.push_back(run_once(P, Q, opt, &score, &T));
      same = same && score == score0 && T == T0;
    }
    std::sort(ms.begin(), ms.end());
    const char* dev = std::getenv("S4PCS_DEVICES");  // candidate sharding over device contexts, set by the caller
    const char* bat = std::getenv("S4PCS_BATCH");    // bases per launch chain (s4g_try_bases), set by the caller
    std::printf("{\"batch\": %s, \"devices\": \"%s\", \"lanes\": %d, \"sample_size\": %d, \"reps\": %d, \"median_ms\": %.2f, \"min_ms\": %.2f, \"max_ms\": %.2f, \"score\": %g, "
                "\"identical_to_lanes1\": %s}\n", bat ? bat : "1", dev ? dev : "1", L, int(opt.sample_size), reps, ms[ms.size() / 2], ms.front(), ms.back(), score,
                same ? "true" : "false");
  }
  return 0;
}
