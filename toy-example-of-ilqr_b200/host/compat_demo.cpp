// Exercises the C++ drop-in (cilqr_solver_compat.hpp) the way src/motion_planning.cpp:178 and
// :194-197 use the reference class: construct from a config object, call solve() per tick, apply
// x.row(1).  Reads a scenario dump written by tests/test_gpu_compat_cpp.py and prints u and x.
//
//   compat_demo <scenario.txt> [ticks]
#include <any>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <string>

#include "cilqr_solver_compat.hpp"

// Same accessor surface as the reference's GlobalConfig (include/global_config.hpp:33-34).
class MiniConfig {
  public:
    std::map<std::string, std::any> m;
    template <typename T>
    T get_config(const std::string& key) const {
        auto it = m.find(key);
        if (it == m.end()) {
            std::cerr << "Configuration key not found: " << key << std::endl;
            return T();
        }
        return std::any_cast<T>(it->second);
    }
};
struct Line {  // ReferenceLine / RoutingLine: three parallel vectors (include/utils.hpp:44-46, :65-67)
    std::vector<double> x, y, yaw;
    size_t size() const { return x.size(); }
};

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s scenario.txt [ticks]\n", argv[0]);
        return 2;
    }
    std::ifstream f(argv[1]);
    if (!f) {
        std::fprintf(stderr, "cannot open %s\n", argv[1]);
        return 2;
    }
    int ticks = argc > 2 ? std::atoi(argv[2]) : 1;
    MiniConfig cfg;
    int n_keys;
    f >> n_keys;
    for (int i = 0; i < n_keys; ++i) {
        std::string key, type, val;
        f >> key >> type >> val;
        if (type == "d") cfg.m[key] = std::stod(val);
        else if (type == "i") cfg.m[key] = std::stoi(val);
        else if (type == "b") cfg.m[key] = bool(std::stoi(val) != 0);
        else cfg.m[key] = val;
    }
    Line ref;
    int M;
    f >> M;
    ref.x.resize(M); ref.y.resize(M); ref.yaw.resize(M);
    for (int i = 0; i < M; ++i) f >> ref.x[i] >> ref.y[i] >> ref.yaw[i];
    int n_obs, T;
    f >> n_obs >> T;
    std::vector<Line> tracks(n_obs);
    for (int j = 0; j < n_obs; ++j) {
        tracks[j].x.resize(T); tracks[j].y.resize(T); tracks[j].yaw.resize(T);
        for (int k = 0; k < T; ++k) f >> tracks[j].x[k] >> tracks[j].y[k] >> tracks[j].yaw[k];
    }
    cilqr_compat::Vector4d ego;
    cilqr_compat::Vector2d borders;
    double target_velocity;
    f >> ego[0] >> ego[1] >> ego[2] >> ego[3] >> target_velocity >> borders[0] >> borders[1];

    try {
        cilqr_compat::CILQRSolver solver(&cfg);
        for (int t = 0; t < ticks; ++t) {
            // utils::get_sub_routing_lines (src/utils.cpp:88-103): tracks from tick t on
            std::vector<Line> sub(n_obs);
            for (int j = 0; j < n_obs; ++j) {
                sub[j].x.assign(tracks[j].x.begin() + t, tracks[j].x.end());
                sub[j].y.assign(tracks[j].y.begin() + t, tracks[j].y.end());
                sub[j].yaw.assign(tracks[j].yaw.begin() + t, tracks[j].yaw.end());
            }
            auto [u, x] = solver.solve(ego, ref, target_velocity, sub, borders);
            std::printf("tick %d status %d iters %d J %.17g %.17g\n", t, int(solver.status()), solver.iterations(),
                        solver.initial_cost(), solver.final_cost());
            for (int i = 0; i < u.rows(); ++i) std::printf("u %.17g %.17g\n", u(i, 0), u(i, 1));
            for (int i = 0; i < x.rows(); ++i) std::printf("x %.17g %.17g %.17g %.17g\n", x(i, 0), x(i, 1), x(i, 2), x(i, 3));
            auto r1 = x.row(1);  // ego_state = new_x.row(1) (motion_planning.cpp:197)
            ego = {r1[0], r1[1], r1[2], r1[3]};
        }
    } catch (const std::out_of_range& e) {
        std::printf("out_of_range %s\n", e.what());
        return 3;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
