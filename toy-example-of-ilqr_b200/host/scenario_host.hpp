// Host-side scenario preparation in C++: everything the reference does
// between reading a scenario file and calling CILQRSolver::solve, without Eigen / yaml-cpp.
//
//   GlobalConfig      same accessor surface as include/global_config.hpp (get_instance(path),
//                     get_config<T>(key), has_key), filled by a reader for the YAML subset the four
//                     shipped scenario files use (nested block maps, scalars, '#' comments, quoted
//                     strings, flow lists, a block list of flow lists); keys and defaults follow
//                     src/global_config.cpp:22-92.
//   CubicSpline(2D)   natural cubic spline through the lane knots (src/cubic_spline.cpp:17-169).
//   ReferenceLine     the spline sampled every 0.1 m with a lateral offset (src/utils.cpp:21-35, :60-67).
//   RoutingLine       (x, y, yaw) per tick (include/utils.hpp:53-68).
//   build_scenario    lanes, borders, constant-speed obstacle tracks along the nearest centre line,
//                     oncoming if yaw0 > pi/2 (src/motion_planning.cpp:91-160; the random noise of
//                     :163-171 is not applied).
#pragma once

#include <algorithm>
#include <any>
#include <cmath>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace cilqr_host {

// ---------------------------------------------------------------------------------------------
// YAML subset
// ---------------------------------------------------------------------------------------------
struct YamlNode {
    enum Kind { Null, Scalar, Map, List } kind = Null;
    std::string scalar;
    std::vector<std::pair<std::string, YamlNode>> map;
    std::vector<YamlNode> list;
    const YamlNode* find(const std::string& k) const {
        for (auto& kv : map)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

namespace detail {
inline std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
inline std::string strip_comment(const std::string& s) {
    bool in_s = false, in_d = false;
    for (size_t i = 0; i < s.size(); ++i) {
        char c = s[i];
        if (c == '\'' && !in_d) in_s = !in_s;
        else if (c == '"' && !in_s) in_d = !in_d;
        else if (c == '#' && !in_s && !in_d && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
    }
    return s;
}
inline std::string unquote(const std::string& s) {
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\'')))
        return s.substr(1, s.size() - 2);
    return s;
}
// "[a, b, [c, d]]" or a plain scalar
inline YamlNode parse_flow(const std::string& text, size_t& pos) {
    while (pos < text.size() && (text[pos] == ' ' || text[pos] == '\t')) ++pos;
    YamlNode n;
    if (pos < text.size() && text[pos] == '[') {
        n.kind = YamlNode::List;
        ++pos;
        for (;;) {
            while (pos < text.size() && (text[pos] == ' ' || text[pos] == ',')) ++pos;
            if (pos >= text.size()) throw std::runtime_error("yaml: unterminated flow list");
            if (text[pos] == ']') {
                ++pos;
                break;
            }
            n.list.push_back(parse_flow(text, pos));
        }
        return n;
    }
    size_t start = pos;
    while (pos < text.size() && text[pos] != ',' && text[pos] != ']') ++pos;
    n.kind = YamlNode::Scalar;
    n.scalar = unquote(trim(text.substr(start, pos - start)));
    return n;
}
struct Line {
    int indent;
    std::string text;
};
inline YamlNode parse_block(const std::vector<Line>& lines, size_t& i, int indent) {
    YamlNode node;
    if (i >= lines.size()) return node;
    if (lines[i].text.rfind("- ", 0) == 0 || lines[i].text == "-") {
        node.kind = YamlNode::List;
        while (i < lines.size() && lines[i].indent == indent && lines[i].text.rfind("-", 0) == 0) {
            std::string rest = trim(lines[i].text.substr(1));
            ++i;
            size_t p = 0;
            node.list.push_back(parse_flow(rest, p));
        }
        return node;
    }
    node.kind = YamlNode::Map;
    while (i < lines.size() && lines[i].indent == indent) {
        const std::string& t = lines[i].text;
        size_t colon = t.find(':');
        if (colon == std::string::npos) throw std::runtime_error("yaml: expected 'key: value' in '" + t + "'");
        std::string key = unquote(trim(t.substr(0, colon)));
        std::string val = trim(t.substr(colon + 1));
        ++i;
        if (!val.empty()) {
            size_t p = 0;
            node.map.emplace_back(key, parse_flow(val, p));
        } else if (i < lines.size() && lines[i].indent > indent) {
            node.map.emplace_back(key, parse_block(lines, i, lines[i].indent));
        } else if (i < lines.size() && lines[i].indent == indent && lines[i].text.rfind("-", 0) == 0) {
            node.map.emplace_back(key, parse_block(lines, i, indent));  // list at the key's own indent
        } else {
            node.map.emplace_back(key, YamlNode());
        }
    }
    return node;
}
}  // namespace detail

inline YamlNode parse_yaml(std::istream& in) {
    std::vector<detail::Line> lines;
    std::string raw;
    while (std::getline(in, raw)) {
        std::string s = detail::strip_comment(raw);
        if (detail::trim(s).empty()) continue;
        int indent = int(s.find_first_not_of(' '));
        lines.push_back({indent, detail::trim(s)});
    }
    size_t i = 0;
    return lines.empty() ? YamlNode() : detail::parse_block(lines, i, lines[0].indent);
}

// ---------------------------------------------------------------------------------------------
// GlobalConfig (include/global_config.hpp surface; src/global_config.cpp:17-131 behaviour)
// ---------------------------------------------------------------------------------------------
class GlobalConfig {
  public:
    static GlobalConfig* get_instance(const std::string& path = "") {
        if (!instance()) {
            if (path.empty()) throw std::runtime_error("GlobalConfig is not initialized!");
            instance() = new GlobalConfig();
            instance()->load_file(path);
        }
        return instance();
    }
    static void destroy_instance() {
        delete instance();
        instance() = nullptr;
    }
    bool has_key(const std::string& key) const { return config_map.find(key) != config_map.end(); }
    template <typename T>
    T get_config(const std::string& key) const {
        auto it = config_map.find(key);
        if (it != config_map.end()) {
            try {
                return std::any_cast<T>(it->second);
            } catch (const std::bad_any_cast&) {
                std::cerr << "Type mismatch for key: " << key << std::endl;
            }
        } else {
            std::cerr << "Configuration key not found: " << key << std::endl;
        }
        return T();
    }
    void load_file(const std::string& path) {
        std::ifstream f(path);
        if (!f) throw std::runtime_error("cannot open " + path);
        load(parse_yaml(f));
    }
    void load(const YamlNode& root) {
        auto sec = [&](const char* s) -> const YamlNode& {
            const YamlNode* n = root.find(s);
            static const YamlNode empty;
            return n ? *n : empty;
        };
        auto num = [&](const YamlNode& m, const char* k, const double* def = nullptr) -> double {
            const YamlNode* n = m.find(k);
            if (!n || n->kind != YamlNode::Scalar) {
                if (def) return *def;
                throw std::runtime_error(std::string("yaml: missing key ") + k);
            }
            return std::stod(n->scalar);
        };
        auto str = [&](const YamlNode& m, const char* k, const char* def) -> std::string {
            const YamlNode* n = m.find(k);
            if (!n || n->kind != YamlNode::Scalar) {
                if (def) return def;
                throw std::runtime_error(std::string("yaml: missing key ") + k);
            }
            return n->scalar;
        };
        auto boolean = [&](const YamlNode& m, const char* k, const bool* def = nullptr) -> bool {
            const YamlNode* n = m.find(k);
            if (!n || n->kind != YamlNode::Scalar) {
                if (def) return *def;
                throw std::runtime_error(std::string("yaml: missing key ") + k);
            }
            return n->scalar == "true" || n->scalar == "True" || n->scalar == "yes";
        };
        auto vec = [&](const YamlNode& n) {
            std::vector<double> v;
            for (auto& e : n.list) v.push_back(std::stod(e.scalar));
            return v;
        };
        config_map["max_simulation_time"] = num(root, "max_simulation_time");
        config_map["delta_t"] = num(root, "delta_t");
        const YamlNode& lqr = sec("lqr");
        for (const char* k : {"N", "nx", "nu"}) config_map[std::string("lqr/") + k] = int(num(lqr, k));
        for (const char* k : {"w_pos", "w_vel", "w_yaw", "w_acc", "w_stl", "obstacle_exp_q1", "obstacle_exp_q2",
                              "state_exp_q1", "state_exp_q2"})
            config_map[std::string("lqr/") + k] = num(lqr, k);
        config_map["lqr/slove_type"] = str(lqr, "slove_type", nullptr);
        const double d_rho = 1.0, d_gamma = 0.0, d_maxrho = 100.0, d_maxmu = 1000.0;  // global_config.cpp:34-37
        config_map["lqr/alm_rho_init"] = num(lqr, "alm_rho_init", &d_rho);
        config_map["lqr/alm_gamma"] = num(lqr, "alm_gamma", &d_gamma);
        config_map["lqr/max_rho"] = num(lqr, "max_rho", &d_maxrho);
        config_map["lqr/max_mu"] = num(lqr, "max_mu", &d_maxmu);
        config_map["lqr/use_last_solution"] = boolean(lqr, "use_last_solution");
        const YamlNode& it = sec("iteration");
        config_map["iteration/max_iter"] = int(num(it, "max_iter"));
        for (const char* k : {"init_lamb", "lamb_decay", "lamb_amplify", "max_lamb", "convergence_threshold",
                              "accept_step_threshold"})
            config_map[std::string("iteration/") + k] = num(it, k);
        const YamlNode& veh = sec("vehicle");
        config_map["vehicle/reference_point"] = str(veh, "reference_point", "gravity_center");  // :54-55
        for (const char* k : {"target_velocity", "wheelbase", "width", "length", "velo_max", "velo_min", "yaw_lim",
                              "acc_max", "acc_min", "stl_lim", "d_safe"})
            config_map[std::string("vehicle/") + k] = num(veh, k);
        const YamlNode& lane = sec("laneline");
        const YamlNode* ref = lane.find("reference");
        if (!ref) throw std::runtime_error("yaml: missing laneline/reference");
        config_map["laneline/reference/x"] = vec(*ref->find("x"));
        config_map["laneline/reference/y"] = vec(*ref->find("y"));
        config_map["laneline/border"] = vec(*lane.find("border"));
        config_map["laneline/center_line"] = vec(*lane.find("center_line"));
        std::vector<std::vector<double>> ic;
        for (auto& row : sec("initial_condition").list) ic.push_back(vec(row));
        config_map["initial_condition"] = ic;
        const YamlNode& vis = sec("visualization");
        const bool no = false;
        config_map["visualization/show_reference_line"] = boolean(vis, "show_reference_line", &no);
        config_map["visualization/show_obstacle_boundary"] = boolean(vis, "show_obstacle_boundary", &no);
        if (vis.find("x_lim")) config_map["visualization/x_lim"] = vec(*vis.find("x_lim"));
        if (vis.find("y_lim")) config_map["visualization/y_lim"] = vec(*vis.find("y_lim"));
    }

    // the flat "section/key" map (what the reference's GlobalConfig holds in its private config_map)
    const std::unordered_map<std::string, std::any>& entries() const { return config_map; }

  private:
    static GlobalConfig*& instance() {
        static GlobalConfig* p = nullptr;
        return p;
    }
    std::unordered_map<std::string, std::any> config_map;
};

// ---------------------------------------------------------------------------------------------
// splines and lines
// ---------------------------------------------------------------------------------------------
class CubicSpline {
  public:
    CubicSpline() {}
    CubicSpline(const std::vector<double>& xs, const std::vector<double>& ys) : x(xs), a(ys) {
        const int n = int(x.size());
        std::vector<double> h(n - 1);
        for (int i = 0; i + 1 < n; ++i) {
            h[i] = x[i + 1] - x[i];
            if (h[i] < 0) throw std::invalid_argument("x coordinates must be sorted in ascending order");
        }
        // natural spline system (src/cubic_spline.cpp:41-69), dense Gaussian elimination with pivoting
        std::vector<std::vector<double>> A(n, std::vector<double>(n, 0.0));
        std::vector<double> rhs(n, 0.0);
        A[0][0] = 1.0;
        for (int i = 0; i + 1 < n; ++i) {
            if (i != n - 2) A[i + 1][i + 1] = 2.0 * (h[i] + h[i + 1]);
            A[i + 1][i] = h[i];
            A[i][i + 1] = h[i];
        }
        A[0][1] = 0.0;
        A[n - 1][n - 2] = 0.0;
        A[n - 1][n - 1] = 1.0;
        for (int i = 0; i + 2 < n; ++i)
            rhs[i + 1] = 3.0 * (a[i + 2] - a[i + 1]) / h[i + 1] - 3.0 * (a[i + 1] - a[i]) / h[i];
        c = solve(A, rhs);
        for (int i = 0; i + 1 < n; ++i) {
            d.push_back((c[i + 1] - c[i]) / (3.0 * h[i]));
            b.push_back((a[i + 1] - a[i]) / h[i] - h[i] * (c[i + 1] + 2 * c[i]) / 3.0);
        }
    }
    double position(double s) const {
        int i = segment(s);
        double dx = s - x[i];
        return a[i] + b[i] * dx + c[i] * std::pow(dx, 2) + d[i] * std::pow(dx, 3);
    }
    double first_derivative(double s) const {
        int i = segment(s);
        double dx = s - x[i];
        return b[i] + 2.0 * c[i] * dx + 3.0 * d[i] * std::pow(dx, 2);
    }

  private:
    int segment(double s) const {
        if (s < x.front() || s > x.back()) throw std::invalid_argument("received value out of the pre-defined range");
        int i = int(std::upper_bound(x.begin(), x.end(), s) - x.begin()) - 1;
        return std::min(i, int(x.size()) - 2);  // s == last knot: the reference reads past its arrays; clamp
    }
    static std::vector<double> solve(std::vector<std::vector<double>> A, std::vector<double> r) {
        const int n = int(r.size());
        for (int k = 0; k < n; ++k) {
            int p = k;
            for (int i = k + 1; i < n; ++i)
                if (std::fabs(A[i][k]) > std::fabs(A[p][k])) p = i;
            std::swap(A[k], A[p]);
            std::swap(r[k], r[p]);
            for (int i = k + 1; i < n; ++i) {
                double f = A[i][k] / A[k][k];
                if (f == 0.0) continue;
                for (int j = k; j < n; ++j) A[i][j] -= f * A[k][j];
                r[i] -= f * r[k];
            }
        }
        for (int k = n - 1; k >= 0; --k) {
            double s = r[k];
            for (int j = k + 1; j < n; ++j) s -= A[k][j] * r[j];
            r[k] = s / A[k][k];
        }
        return r;
    }
    std::vector<double> x, a, b, c, d;
};

class CubicSpline2D {
  public:
    CubicSpline2D() {}
    CubicSpline2D(const std::vector<double>& xs, const std::vector<double>& ys) {
        s.push_back(0.0);
        for (size_t i = 1; i < xs.size(); ++i) s.push_back(s.back() + std::hypot(xs[i] - xs[i - 1], ys[i] - ys[i - 1]));
        sx = CubicSpline(s, xs);
        sy = CubicSpline(s, ys);
    }
    void position(double t, double& px, double& py) const {
        px = sx.position(t);
        py = sy.position(t);
    }
    double yaw(double t) const { return std::atan2(sy.first_derivative(t), sx.first_derivative(t)); }
    std::vector<double> s;

  private:
    CubicSpline sx, sy;
};

struct ReferenceLine {
    ReferenceLine(const std::vector<double>& xs, const std::vector<double>& ys, double width = 0, double accuracy = 0.1)
        : spline(xs, ys), delta_s(accuracy), delta_d(width) {
        for (double t = 0.0; t <= spline.s.back(); t += delta_s) {
            double px, py, lyaw = spline.yaw(t);
            spline.position(t, px, py);
            x.push_back(px - width * std::sin(lyaw));
            y.push_back(py + width * std::cos(lyaw));
            yaw.push_back(lyaw);
            longitude.push_back(t);
        }
    }
    void calc_position(double t, double out[3]) const {
        double px, py, lyaw = spline.yaw(t);
        spline.position(t, px, py);
        out[0] = px - delta_d * std::sin(lyaw);
        out[1] = py + delta_d * std::cos(lyaw);
        out[2] = lyaw;
    }
    size_t size() const { return x.size(); }
    double length() const { return spline.s.back(); }
    std::vector<double> x, y, yaw, longitude;
    CubicSpline2D spline;
    double delta_s, delta_d;
};

struct RoutingLine {
    std::vector<double> x, y, yaw;
};

struct Scenario {
    std::vector<ReferenceLine> borders, center_lines;
    double road_borders[2];
    std::vector<RoutingLine> routing_lines;  // [0] = ego, 1.. = obstacles
    std::vector<std::vector<double>> initial_conditions;
    double delta_t, max_simulation_time, target_velocity;
};

// src/motion_planning.cpp:91-173 without the random noise
inline Scenario build_scenario(const GlobalConfig& cfg) {
    Scenario sc;
    sc.delta_t = cfg.get_config<double>("delta_t");
    sc.max_simulation_time = cfg.get_config<double>("max_simulation_time");
    sc.target_velocity = cfg.get_config<double>("vehicle/target_velocity");
    auto rx = cfg.get_config<std::vector<double>>("laneline/reference/x");
    auto ry = cfg.get_config<std::vector<double>>("laneline/reference/y");
    auto border_w = cfg.get_config<std::vector<double>>("laneline/border");
    auto center_w = cfg.get_config<std::vector<double>>("laneline/center_line");
    sc.initial_conditions = cfg.get_config<std::vector<std::vector<double>>>("initial_condition");
    for (double w : border_w) sc.borders.emplace_back(rx, ry, w);
    for (double w : center_w) sc.center_lines.emplace_back(rx, ry, w);
    std::sort(border_w.begin(), border_w.end(), std::greater<double>());
    sc.road_borders[0] = border_w.front();
    sc.road_borders[1] = border_w.back();
    const size_t n = sc.initial_conditions.size();
    sc.routing_lines.resize(n);
    for (size_t idx = 0; idx < n; ++idx) {
        const auto& ic = sc.initial_conditions[idx];
        size_t line_num = 0;
        double start_s = sc.center_lines[0].length(), min_diff = -1.0;
        for (size_t l = 0; l < sc.center_lines.size(); ++l) {
            const ReferenceLine& cl = sc.center_lines[l];
            for (size_t i = 1; i < cl.size(); ++i) {
                double last = std::hypot(cl.x[i - 1] - ic[0], cl.y[i - 1] - ic[1]);
                double cur = std::hypot(cl.x[i] - ic[0], cl.y[i] - ic[1]);
                if (cur > last) {
                    if (min_diff < 0 || last < min_diff) {
                        min_diff = last;
                        line_num = l;
                        start_s = cl.longitude[i - 1];
                    }
                    break;
                }
            }
        }
        const ReferenceLine& cl = sc.center_lines[line_num];
        for (double t = 0.0; t < sc.max_simulation_time + 10; t += sc.delta_t) {
            double pos[3];
            if (ic[3] <= M_PI_2) {
                double s = std::min(start_s + t * ic[2], cl.longitude.back());
                cl.calc_position(s, pos);
            } else {
                double s = std::max(start_s - t * ic[2], cl.longitude.front());
                cl.calc_position(s, pos);
                pos[2] = std::fmod(pos[2] + M_PI, 2 * M_PI);
            }
            sc.routing_lines[idx].x.push_back(pos[0]);
            sc.routing_lines[idx].y.push_back(pos[1]);
            sc.routing_lines[idx].yaw.push_back(pos[2]);
        }
    }
    return sc;
}

}  // namespace cilqr_host
