// include/cilqr_solver.hpp — B200 build.  TEST INFRASTRUCTURE: the stub INTEGRATION.md section 1 tells a
// reference maintainer to put in place of the reference's include/cilqr_solver.hpp, so that the reference's own,
// unmodified src/motion_planning.cpp (lines 178 and 194-197) drives libcilqr_b200.so.
#pragma once
#define CILQR_COMPAT_USE_EIGEN            // Eigen::Vector4d / MatrixX2d / MatrixX4d in the signatures
#include "utils.hpp"                      // ReferenceLine, RoutingLine (x, y, yaw vectors)
#include "global_config.hpp"              // GlobalConfig::get_config<T>(key)
#include <cilqr_solver_compat.hpp>
using CILQRSolver = cilqr_compat::CILQRSolver;
using cilqr_compat::LQRSolveStatus;
