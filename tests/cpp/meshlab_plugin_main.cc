// Test driver (ours) for the REFERENCE's MeshLab plugin GlobalRegistrationPlugin, whose two source files are used
// unchanged from <reference>/demos/MeshlabPlugin/filter_globalregistration and compiled against the product's headers with
// the MeshLab/Qt stub of tests/stubs/meshlab.  Usage: meshlab_plugin_test ref.xyz target.xyz overlap_percent delta n_samples
#include "globalregistration.cpp"  // the reference's plugin, found through -I

#include <cstdlib>
#include <fstream>

static bool load(const char* path, CMeshO& m) {
  std::ifstream f(path);
  CVertexO v;
  while (f >> v.p.v[0] >> v.p.v[1] >> v.p.v[2]) m.vert.push_back(v);
  return !m.vert.empty();
}

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  MeshModel ref, target;
  if (!load(argv[1], ref.cm) || !load(argv[2], target.cm)) return 3;
  MeshDocument doc;
  doc.current = &ref;
  GlobalRegistrationPlugin plugin;
  QAction* action = plugin.actionList.at(0);
  std::printf("filter: %s (%s)\n", plugin.filterName(plugin.ID(action)).str().c_str(), plugin.pluginName().str().c_str());
  RichParameterSet par;
  plugin.initParameterSet(action, doc, par);             // the plugin's own defaults ...
  par.at("refMesh").mesh = &ref;                         // ... then what a user would set in the dialog
  par.at("targetMesh").mesh = &target;
  par.at("overlap").number = std::atof(argv[3]);
  par.at("delta").number = std::atof(argv[4]);
  par.at("nbSamples").number = std::atoi(argv[5]);
  par.at("max_time_seconds").number = 1000;
  if (!plugin.applyFilter(action, doc, par, nullptr)) return 4;
  for (int r = 0; r < 4; ++r)
    std::printf("Tr-row: %.9g %.9g %.9g %.9g\n", target.cm.Tr.m[r][0], target.cm.Tr.m[r][1], target.cm.Tr.m[r][2], target.cm.Tr.m[r][3]);
  return 0;
}
