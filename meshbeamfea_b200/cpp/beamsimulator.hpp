// beamsimulator.hpp -- C++17 drop-in for the reference's simulation API
// [REF src/beamsimulator.hpp:29-52,78-95], implemented on libmbfea.so
// (include/mbfea.h).  Same type names, members, method signatures and error
// behaviour as the reference header, so src/gui.cpp, src/drawer.cpp,
// src/colorbar.cpp and src/beamsimulatortester.cpp compile against it unchanged:
//
//   struct NodalForce / SimulationJob / SimulationResults     (same members)
//   class  BeamSimulator { setSimulation, executeSimulation, getResults,
//                          setResultsNodalDisplacementCSVFilepath,
//                          setResultsElementalForcesCSVFilepath }
//
// Where <Eigen/Core> is available (the reference's build) the members are the
// Eigen types of the reference.  Without Eigen (this repository's CI image has
// none) a minimal column-major stand-in with the same accessors is used, so the
// header can be compiled and exercised here.
//
// Behavioural notes:
//  * printInfo()'s O(V+E) stdout dump [REF src/beamsimulator.cpp:28-38] is not
//    reproduced.
//  * CSV files are written when a file name is set, like the reference
//    [REF src/beamsimulator.cpp:47-54]; the default names of
//    [REF src/beamsimulator.hpp:90-91] are kept but can be cleared with "".
//  * GSL Expects() violations terminate in the reference; here they throw
//    mbfea::ContractViolation (derived from std::logic_error), which terminates
//    the same way when uncaught.
#ifndef BEAMSIMULATOR_HPP
#define BEAMSIMULATOR_HPP

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mbfea.h"

#if defined(MBFEA_USE_EIGEN) || (__has_include(<Eigen/Core>) && !defined(MBFEA_NO_EIGEN))
#include <Eigen/Core>
#define MBFEA_HAVE_EIGEN 1
#else
#define MBFEA_HAVE_EIGEN 0
#endif

namespace mbfea {
struct ContractViolation : std::logic_error { using std::logic_error::logic_error; };
inline void expects(bool ok, const char *what) { if (!ok) throw ContractViolation(what); }

#if !MBFEA_HAVE_EIGEN
// Minimal column-major dense matrix with the accessors the reference uses on
// its Eigen members: rows(), cols(), operator()(i,j), data(), resize(), row-less.
template <class T, int Cols = -1>
class Matrix {
public:
    Matrix() = default;
    Matrix(std::ptrdiff_t r, std::ptrdiff_t c) { resize(r, c); }
    explicit Matrix(std::ptrdiff_t r) { resize(r, Cols > 0 ? Cols : 1); }
    void resize(std::ptrdiff_t r, std::ptrdiff_t c) { r_ = r; c_ = c; d_.assign((size_t)(r * c), T()); }
    void resize(std::ptrdiff_t r) { resize(r, Cols > 0 ? Cols : 1); }
    std::ptrdiff_t rows() const { return r_; }
    std::ptrdiff_t cols() const { return c_; }
    std::ptrdiff_t size() const { return r_ * c_; }
    T &operator()(std::ptrdiff_t i, std::ptrdiff_t j) { return d_[(size_t)(j * r_ + i)]; }
    const T &operator()(std::ptrdiff_t i, std::ptrdiff_t j) const { return d_[(size_t)(j * r_ + i)]; }
    T &operator()(std::ptrdiff_t i) { return d_[(size_t)i]; }
    const T &operator()(std::ptrdiff_t i) const { return d_[(size_t)i]; }
    T *data() { return d_.data(); }
    const T *data() const { return d_.data(); }
private:
    std::ptrdiff_t r_ = 0, c_ = Cols > 0 ? Cols : 0;
    std::vector<T> d_;
};
#endif
}  // namespace mbfea

#if MBFEA_HAVE_EIGEN
namespace mbfea_types {
using MatrixX3d = Eigen::MatrixX3d; using MatrixX2i = Eigen::MatrixX2i; using MatrixXd = Eigen::MatrixXd;
using VectorXi = Eigen::VectorXi; using VectorXd = Eigen::VectorXd;
}
#else
namespace Eigen {  // stand-ins under the reference's spelling
using MatrixX3d = mbfea::Matrix<double, 3>; using MatrixX2i = mbfea::Matrix<int, 2>;
using MatrixXd = mbfea::Matrix<double>; using VectorXi = mbfea::Matrix<int, 1>; using VectorXd = mbfea::Matrix<double, 1>;
}
#endif

#ifndef BEAM_HPP  // the reference's src/beam.hpp, when included first, wins
#define BEAM_HPP
struct BeamDimensions {  // [REF src/beam.hpp:9-17]
    float b;
    float h;
    BeamDimensions(const float &width, const float &height) : b(width), h(height)
    {
        mbfea::expects(width > 0 && height > 0, "BeamDimensions: width > 0 && height > 0");
    }
    BeamDimensions() : b(1), h(1) {}
};
struct BeamMaterial {  // [REF src/beam.hpp:19-28]
    float poissonsRatio;
    float youngsModulusGPascal;
    BeamMaterial(const float &nu, const float &E) : poissonsRatio(nu), youngsModulusGPascal(E)
    {
        mbfea::expects(nu <= 0.5 && nu >= -1, "BeamMaterial: poissonsRatio <= 0.5 && poissonsRatio >= -1");
    }
    BeamMaterial() : poissonsRatio(0.3f), youngsModulusGPascal(210) {}
};
#endif

struct NodalForce {  // [REF src/beamsimulator.hpp:29-33]; gsl::index == std::ptrdiff_t
    std::ptrdiff_t index;
    std::ptrdiff_t dof;
    double magnitude;
};

struct SimulationJob {  // [REF src/beamsimulator.hpp:35-43]
    Eigen::MatrixX3d nodes;
    Eigen::MatrixX2i elements;
    Eigen::MatrixX3d elementalNormals;
    Eigen::VectorXi fixedNodes;
    std::vector<NodalForce> nodalForces;
    std::vector<BeamDimensions> beamDimensions;
    std::vector<BeamMaterial> beamMaterial;
};

struct SimulationResults {  // [REF src/beamsimulator.hpp:45-52]
    std::vector<Eigen::VectorXd> edgeForces;  ///< #force components x 2*#edges
    Eigen::MatrixXd nodalDisplacements;       ///< #nodes x 6
    SimulationResults() {}
};

class BeamSimulator {  // [REF src/beamsimulator.hpp:78-95]
public:
    BeamSimulator() { open(nullptr); }
    explicit BeamSimulator(const mbfea_options &opts) { open(&opts); }
    ~BeamSimulator() { mbfea_destroy(ctx_); }
    BeamSimulator(const BeamSimulator &) = delete;
    BeamSimulator &operator=(const BeamSimulator &) = delete;

    void setSimulation(const SimulationJob &job)
    {
        haveJob_ = false;
        const std::ptrdiff_t V = job.nodes.rows(), E = job.elements.rows();
        // Expects(...) of FeaSimulationJob::setElements [REF src/beamsimulator.cpp:73-75]
        mbfea::expects(E == job.elementalNormals.rows(), "elements.rows() == elementalNormals.rows()");
        mbfea::expects((size_t)E == job.beamDimensions.size(), "elements.rows() == beamDimensions.size()");
        mbfea::expects((size_t)E == job.beamMaterial.size(), "elements.rows() == beamMaterial.size()");
        static_assert(sizeof(BeamDimensions) == 2 * sizeof(float) && sizeof(BeamMaterial) == 2 * sizeof(float),
                      "section PODs are passed as raw {float,float} pairs");
        std::vector<int64_t> fn(job.nodalForces.size()), fd(job.nodalForces.size());
        std::vector<double> fm(job.nodalForces.size());
        for (size_t i = 0; i < job.nodalForces.size(); ++i) {
            fn[i] = job.nodalForces[i].index; fd[i] = job.nodalForces[i].dof; fm[i] = job.nodalForces[i].magnitude;
        }
        std::vector<int32_t> el((size_t)(2 * E)), fx((size_t)job.fixedNodes.size());
        for (std::ptrdiff_t i = 0; i < 2 * E; ++i) el[(size_t)i] = job.elements.data()[i];
        for (std::ptrdiff_t i = 0; i < job.fixedNodes.size(); ++i) fx[(size_t)i] = job.fixedNodes.data()[i];
        const int rc = mbfea_set_job(ctx_, V, E, job.nodes.data(), el.data(), job.elementalNormals.data(),
                                     reinterpret_cast<const float *>(job.beamDimensions.data()),
                                     reinterpret_cast<const float *>(job.beamMaterial.data()), (int64_t)fx.size(),
                                     fx.data(), (int64_t)fn.size(), fn.data(), fd.data(), fm.data());
        raise(rc);
        V_ = V; E_ = E; haveJob_ = true;
    }

    SimulationResults executeSimulation()
    {
        mbfea::expects(haveJob_, "!simulationJob.isEmpty()");  // [REF src/beamsimulator.cpp:41]
        raise(mbfea_execute(ctx_));
        if (!elementalForcesOutputFilepath.empty())
            raise(mbfea_write_element_forces_csv(ctx_, elementalForcesOutputFilepath.c_str()));
        if (!nodalDisplacementsOutputFilepath.empty())
            raise(mbfea_write_displacements_csv(ctx_, nodalDisplacementsOutputFilepath.c_str()));
        SimulationResults r;
        r.nodalDisplacements.resize(V_, 6);  // column-major V x 6 == the C ABI image
        raise(mbfea_get_displacements(ctx_, r.nodalDisplacements.data()));
        std::vector<double> F((size_t)(12 * E_));
        raise(mbfea_get_edge_forces(ctx_, F.data()));
        r.edgeForces.resize(6);
        for (int c = 0; c < 6; ++c) {
            r.edgeForces[(size_t)c].resize(2 * E_);
            for (std::ptrdiff_t i = 0; i < 2 * E_; ++i) r.edgeForces[(size_t)c](i) = F[(size_t)(c * 2 * E_ + i)];
        }
        results = r;
        return results;
    }

    SimulationResults getResults() const { return results; }
    void setResultsNodalDisplacementCSVFilepath(const std::string &outputPath) { nodalDisplacementsOutputFilepath = outputPath; }
    void setResultsElementalForcesCSVFilepath(const std::string &outputPath) { elementalForcesOutputFilepath = outputPath; }

    // additions (not in the reference): solver statistics and the raw context
    mbfea_stats getStats() const { mbfea_stats s; mbfea_get_stats(ctx_, &s); return s; }
    mbfea_ctx *context() { return ctx_; }

private:
    void open(const mbfea_options *o)
    {
        const int rc = mbfea_create(&ctx_, o);
        if (rc != MBFEA_OK) throw std::runtime_error(std::string("libmbfea: ") + mbfea_last_error(nullptr));
    }
    void raise(int rc) const
    {
        if (rc == MBFEA_OK) return;
        const std::string msg = mbfea_last_error(ctx_);
        switch (rc) {
            case MBFEA_EPRECOND: throw mbfea::ContractViolation(msg);
            case MBFEA_ERANGE_FIXED: throw std::runtime_error(msg);  // [REF src/beamsimulator.cpp:134-137]
            case MBFEA_ERANGE_FORCE: throw std::range_error(msg);    // [REF src/beamsimulator.cpp:164-166]
            default: throw std::runtime_error(std::string(mbfea_status_string(rc)) + ": " + msg);
        }
    }
    mbfea_ctx *ctx_ = nullptr;
    bool haveJob_ = false;
    std::ptrdiff_t V_ = 0, E_ = 0;
    std::string nodalDisplacementsOutputFilepath{"nodal_displacement.csv"};  // [REF src/beamsimulator.hpp:90-91]
    std::string elementalForcesOutputFilepath{"elemental_forces.csv"};
    SimulationResults results;
};

#endif  // BEAMSIMULATOR_HPP
