// Headless replay of the reference's simulator loop (src/motion_planning.cpp:29-197 without the
// plotting): read a scenario file in the reference's YAML format, build lanes / borders / obstacle
// tracks on the host, then per tick call CILQRSolver::solve (the C++ drop-in over libcilqr_b200.so)
// and apply the first step.  Writes one CSV row per tick.
//
//   headless_planner -c scenario.yaml [-t ticks] [-o out.csv]
//   headless_planner -c scenario.yaml -d          dump the prepared scenario arrays (no GPU needed)
#include <unistd.h>

#include <cstdio>
#include <cstdlib>

#include "cilqr_solver_compat.hpp"
#include "scenario_host.hpp"

using cilqr_host::GlobalConfig;
using cilqr_host::RoutingLine;

int main(int argc, char** argv) {
    std::string config_path, out_path;
    int max_ticks = -1, opt;
    bool dump = false;
    while ((opt = getopt(argc, argv, "c:t:o:d")) != -1) {
        switch (opt) {
            case 'c': config_path = optarg; break;
            case 't': max_ticks = std::atoi(optarg); break;
            case 'o': out_path = optarg; break;
            case 'd': dump = true; break;
            default: std::fprintf(stderr, "Usage: %s -c scenario.yaml [-t ticks] [-o out.csv] [-d]\n", argv[0]); return 2;
        }
    }
    if (config_path.empty()) {
        std::fprintf(stderr, "Usage: %s -c scenario.yaml [-t ticks] [-o out.csv] [-d]\n", argv[0]);
        return 2;
    }
    try {
        GlobalConfig* config = GlobalConfig::get_instance(config_path);
        cilqr_host::Scenario sc = cilqr_host::build_scenario(*config);
        const cilqr_host::ReferenceLine& ref = sc.center_lines[0];  // motion_planning.cpp:195
        if (dump) {
            std::printf("ref %zu\n", ref.size());
            for (size_t i = 0; i < ref.size(); ++i) std::printf("%.17g %.17g %.17g\n", ref.x[i], ref.y[i], ref.yaw[i]);
            std::printf("borders %.17g %.17g\n", sc.road_borders[0], sc.road_borders[1]);
            std::printf("tracks %zu %zu\n", sc.routing_lines.size(), sc.routing_lines[0].x.size());
            for (auto& r : sc.routing_lines)
                for (size_t k = 0; k < r.x.size(); ++k) std::printf("%.17g %.17g %.17g\n", r.x[k], r.y[k], r.yaw[k]);
            std::printf("N %d reference_point %s slove_type %s\n", config->get_config<int>("lqr/N"),
                        config->get_config<std::string>("vehicle/reference_point").c_str(),
                        config->get_config<std::string>("lqr/slove_type").c_str());
            return 0;
        }
        FILE* out = out_path.empty() ? stdout : std::fopen(out_path.c_str(), "w");
        if (!out) throw std::runtime_error("cannot open " + out_path);
        std::vector<RoutingLine> obs_prediction(sc.routing_lines.begin() + 1, sc.routing_lines.end());
        cilqr_compat::Vector4d ego = {sc.initial_conditions[0][0], sc.initial_conditions[0][1],
                                      sc.initial_conditions[0][2], sc.initial_conditions[0][3]};
        cilqr_compat::Vector2d borders = {sc.road_borders[0], sc.road_borders[1]};
        cilqr_compat::CILQRSolver solver(config, 0, int(std::max<size_t>(obs_prediction.size(), 1)));
        std::fprintf(out, "t,x,y,v,yaw,acc,steer,iters,status,cost\n");
        int tick = 0;
        for (double t = 0.; t < sc.max_simulation_time; t += sc.delta_t, ++tick) {
            if (max_ticks >= 0 && tick >= max_ticks) break;
            size_t index = size_t(t / sc.delta_t);  // motion_planning.cpp:181
            std::vector<RoutingLine> sub(obs_prediction.size());  // utils::get_sub_routing_lines
            for (size_t j = 0; j < sub.size(); ++j) {
                sub[j].x.assign(obs_prediction[j].x.begin() + index, obs_prediction[j].x.end());
                sub[j].y.assign(obs_prediction[j].y.begin() + index, obs_prediction[j].y.end());
                sub[j].yaw.assign(obs_prediction[j].yaw.begin() + index, obs_prediction[j].yaw.end());
            }
            auto [u, x] = solver.solve(ego, ref, sc.target_velocity, sub, borders);
            std::fprintf(out, "%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g,%d,%d,%.17g\n", t, ego[0], ego[1], ego[2],
                         ego[3], u(0, 0), u(0, 1), solver.iterations(), int(solver.status()), solver.final_cost());
            auto r1 = x.row(1);
            ego = {r1[0], r1[1], r1[2], r1[3]};
        }
        if (out != stdout) std::fclose(out);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "headless_planner: %s\n", e.what());
        return 1;
    }
    return 0;
}
