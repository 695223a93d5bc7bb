// lethe-particles-b200 — the reference's `lethe-particles case.prm` entry point
// (applications/lethe-particles/dem.cc) with the DEM time step on a B200.
//   usage: lethe-particles-b200 case.prm [--device N] [--quiet] [--dump-config]
//   --dump-config prints the lethe_dem_config derived from the .prm as JSON and exits (no GPU needed).
// Errors surface like the reference's: message on stderr, exit code 1.
#include <cstring>
#include <iostream>
#include <sstream>

#include "dem_solver.h"

int main(int argc, char **argv)
{
  try
    {
      if (argc < 2)
        {
          std::cerr << "usage: " << argv[0] << " case.prm [--device N] [--quiet]\n";
          return 1;
        }
      int device = 0;
      bool quiet = false, dump = false;
      for (int a = 2; a < argc; ++a)
        {
          if (!std::strcmp(argv[a], "--device") && a + 1 < argc)
            device = std::atoi(argv[++a]);
          else if (!std::strcmp(argv[a], "--quiet"))
            quiet = true;
          else if (!std::strcmp(argv[a], "--dump-config"))
            dump = true;
        }
      const lethe_b200::DEMParameters prm = lethe_b200::DEMParameters::from_prm_file(argv[1]);
      if (dump)
        {
          const lethe_dem_config c = prm.to_config();
          std::cout.precision(17);
          std::cout << "{\"pp_model\": " << c.pp_model << ", \"pw_model\": " << c.pw_model << ", \"rolling_model\": " << c.rolling_model
                    << ", \"detection\": " << c.detection << ", \"contact_detection_frequency\": " << c.contact_detection_frequency
                    << ", \"cell_order\": " << c.cell_order << ", \"dt\": " << c.dt << ", \"g\": [" << c.g[0] << ", " << c.g[1] << ", "
                    << c.g[2] << "], \"neighborhood_threshold\": " << c.neighborhood_threshold << ", \"d_max\": " << c.d_max
                    << ", \"smallest_contact_search_criterion\": " << c.smallest_contact_search_criterion
                    << ", \"dmt_cut_off_threshold\": " << c.dmt_cut_off_threshold << ", \"f_coefficient_epsd\": " << c.f_coefficient_epsd
                    << ", \"n_types\": " << c.n_types << ", \"restart\": " << c.restart << ", \"young\": [" << c.young[0] << ", "
                    << c.young[1] << "], \"poisson\": [" << c.poisson[0] << "], \"restitution\": [" << c.restitution[0]
                    << "], \"friction\": [" << c.friction[0] << "], \"rolling_friction\": [" << c.rolling_friction[0]
                    << "], \"young_wall\": " << c.young_wall << ", \"friction_wall\": " << c.friction_wall << ", \"grid_lo\": ["
                    << c.grid_lo[0] << ", " << c.grid_lo[1] << ", " << c.grid_lo[2] << "], \"cell_size\": [" << c.cell_size[0] << ", "
                    << c.cell_size[1] << ", " << c.cell_size[2] << "], \"grid_n\": [" << c.grid_n[0] << ", " << c.grid_n[1] << ", "
                    << c.grid_n[2] << "], \"periodic\": [" << c.periodic[0] << ", " << c.periodic[1] << ", " << c.periodic[2]
                    << "], \"n_wall_faces\": "
                    << lethe_b200::box_wall_faces(prm.mesh, prm.outlet_boundaries(), prm.periodic_directions()).size() << "}\n";
          return 0;
        }
      std::ostringstream sink;
      lethe_b200::DEMSolverB200 solver(prm, device, quiet ? static_cast<std::ostream &>(sink) : std::cout);
      solver.solve();
      if (prm.test_enabled)
        solver.print_xyz(std::cout);
    }
  catch (const std::exception &exc)
    {
      std::cerr << "\n----------------------------------------------------\n"
                << "Exception on processing: \n"
                << exc.what() << "\nAborting!\n"
                << "----------------------------------------------------\n";
      return 1;
    }
  return 0;
}
