// `_flexible_robot_pb`: a compiled pybind11 module with the class the reference binds (raisim_gym.cpp:14-47: FlexibleGymEnv over
// VectorizedEnvironment<ENVIRONMENT>), here over the C ABI of libirrl_b200.so (include/irrl_b200.h).  Same constructor
// (resourceDir, cfg) and the same method names and argument lists; a maintainer of the reference who wants the
// module to be called `_flexible_robot` builds this file with -DIRRL_PYMODULE_NAME=_flexible_robot (option B).
//
// Array arguments follow pybind11's Eigen::Ref rules for writable references (RaisimGymEnv.hpp:46-49): float32 (bool for
// `done`), C-contiguous row-major, exact shape, writable -- anything else is a TypeError, never a silent copy.  The call runs
// with the GIL held, like the reference's (no gil_scoped_release in raisim_gym.cpp).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/irrl_b200.h"

namespace py = pybind11;

#ifndef IRRL_PYMODULE_NAME
#define IRRL_PYMODULE_NAME _flexible_robot_pb
#endif

namespace {

void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + irrl_last_error());
}

// pointer of a numpy argument that must already be exactly what the native side writes into
template <typename T> T* ref(py::handle h, std::initializer_list<py::ssize_t> shape, const char* name, bool writable = true) {
    if (!py::isinstance<py::array>(h)) throw py::type_error(std::string(name) + ": expected numpy.ndarray");
    py::array a = py::reinterpret_borrow<py::array>(h);
    const bool is_bool = std::is_same<T, uint8_t>::value;
    const bool dtype_ok = is_bool ? (a.dtype().kind() == 'b' || (a.dtype().kind() == 'u' && a.itemsize() == 1)) : a.dtype().is(py::dtype::of<T>());
    bool shape_ok = a.ndim() == (py::ssize_t)shape.size();
    if (shape_ok) { py::ssize_t i = 0; for (auto s : shape) shape_ok = shape_ok && a.shape(i++) == s; }
    const bool contiguous = (a.flags() & py::array::c_style) != 0;
    if (!dtype_ok || !shape_ok || !contiguous || (writable && !a.writeable())) {
        std::string want = std::string(is_bool ? "bool" : "float32") + "[";
        for (auto s : shape) want += std::to_string(s) + ",";
        want.back() = ']';
        throw py::type_error(std::string(name) + ": incompatible function arguments: expected a writable C-contiguous numpy array " + want +
                             " (no implicit copies, as with Eigen::Ref)");
    }
    return static_cast<T*>(a.mutable_data());
}

class FlexibleGymEnv {
public:
    FlexibleGymEnv(const std::string& resource_dir, const std::string& cfg, int device, int env_offset) {
        check(irrl_create(resource_dir.c_str(), cfg.c_str(), device, env_offset, &h_), "FlexibleGymEnv");
        n_ = irrl_get_num_envs(h_);
    }
    ~FlexibleGymEnv() { if (h_) irrl_destroy(h_); }
    FlexibleGymEnv(const FlexibleGymEnv&) = delete;
    FlexibleGymEnv& operator=(const FlexibleGymEnv&) = delete;

    void init() { check(irrl_init(h_), "init"); }                                                                   // VEC:145
    std::vector<std::string> getExtraInfoNames() {                                                                    // VEC:196
        std::vector<std::string> out; for (int i = 0; i < irrl_get_extra_info_dim(h_); ++i) out.emplace_back(irrl_get_extra_info_name(h_, i)); return out;
    }
    void reset(py::handle ob) { check(irrl_reset(h_, ref<float>(ob, {n_, 35}, "ob")), "reset"); }                     // VEC:201
    void observe(py::handle ob) { check(irrl_observe(h_, ref<float>(ob, {n_, 35}, "ob")), "observe"); }               // VEC:209
    void step(py::handle action, py::handle ob, py::handle reward, py::handle done, py::handle extra) {               // VEC:268
        check(irrl_step(h_, ref<float>(action, {n_, 12}, "action", false), ref<float>(ob, {n_, 35}, "ob"), ref<float>(reward, {n_}, "reward"),
                        ref<uint8_t>(done, {n_}, "done"), ref<float>(extra, {n_, 6}, "extraInfo")), "step");
    }
    void testStep(py::handle action, py::handle ob, py::handle reward, py::handle done, py::handle extra) {           // VEC:280
        check(irrl_test_step(h_, ref<float>(action, {n_, 12}, "action", false), ref<float>(ob, {n_, 35}, "ob"), ref<float>(reward, {n_}, "reward"),
                             ref<uint8_t>(done, {n_}, "done"), ref<float>(extra, {n_, 6}, "extraInfo")), "testStep");
    }
    void setSeed(int seed) { check(irrl_set_seed(h_, seed), "setSeed"); }                                             // VEC:308
    void close() { check(irrl_close(h_), "close"); }                                                                  // VEC:314
    void isTerminalState(py::handle t) { check(irrl_is_terminal_state(h_, ref<uint8_t>(t, {n_}, "terminalState")), "isTerminalState"); }   // VEC:319
    void setSimulationTimeStep(double dt) { check(irrl_set_simulation_time_step(h_, dt), "setSimulationTimeStep"); }  // VEC:326
    void setControlTimeStep(double dt) { check(irrl_set_control_time_step(h_, dt), "setControlTimeStep"); }           // VEC:331
    int getObDim() { return irrl_get_ob_dim(h_); }                                                                    // VEC:336-342
    int getActionDim() { return irrl_get_action_dim(h_); }
    int getExtraInfoDim() { return irrl_get_extra_info_dim(h_); }
    int getNumOfEnvs() { return irrl_get_num_envs(h_); }
    void startRecordingVideo(const std::string& f) { check(irrl_start_recording_video(h_, f.c_str()), "startRecordingVideo"); }   // VEC:292-306
    void stopRecordingVideo() { check(irrl_stop_recording_video(h_), "stopRecordingVideo"); }
    void showWindow() { check(irrl_show_window(h_), "showWindow"); }
    void hideWindow() { check(irrl_hide_window(h_), "hideWindow"); }
    void curriculumUpdate() { check(irrl_curriculum_update(h_), "curriculumUpdate"); }                                // VEC:345
    void OriginState(py::handle o) { check(irrl_origin_state(h_, ref<float>(o, {n_, 41}, "ob")), "OriginState"); }    // VEC:214
    int GetOriginStateDim() { return irrl_get_origin_state_dim(h_); }
    void ReferenceState(py::handle o) { check(irrl_reference_state(h_, ref<float>(o, {n_, 24}, "refer")), "ReferenceState"); }            // VEC:223
    void GetJointEffort(py::handle o) { check(irrl_get_joint_effort(h_, ref<float>(o, {n_, 12}, "joint_effort")), "GetJointEffort"); }    // VEC:231
    void GetGeneralizedForce(py::handle o) { check(irrl_get_generalized_force(h_, ref<float>(o, {n_, 18}, "generalized_force")), "GetGeneralizedForce"); }
    void GetInverseMassMatrix(py::handle o) { check(irrl_get_inverse_mass_matrix(h_, ref<float>(o, {n_, 324}, "inverse_mass")), "GetInverseMassMatrix"); }
    void GetNonlinear(py::handle o) { check(irrl_get_nonlinear(h_, ref<float>(o, {n_, 18}, "nonlinear")), "GetNonlinear"); }
    void SetContactCoefficient(py::handle c) { check(irrl_set_contact_coefficient(h_, ref<float>(c, {n_, 3}, "contact_coeff", false)), "SetContactCoefficient"); }
    void GetSphereInfo(py::handle o) { check(irrl_get_sphere_info(h_, ref<float>(o, {n_, 4}, "sphere_info")), "GetSphereInfo"); }
    // additions of this implementation used by the VecEnv adapter (episode statistics kept on the device)
    void lastEpisodeStats(py::handle r, py::handle l) {
        py::array_t<int32_t, py::array::c_style> la = py::reinterpret_borrow<py::array>(l);
        check(irrl_last_episode_stats(h_, ref<float>(r, {n_}, "ep_return"), la.mutable_data()), "lastEpisodeStats");
    }
    void runningEpisodeStats(py::handle r, py::handle l, bool clear) {
        py::array_t<int32_t, py::array::c_style> la = py::reinterpret_borrow<py::array>(l);
        check(irrl_running_episode_stats(h_, ref<float>(r, {n_}, "ep_return"), la.mutable_data(), clear ? 1 : 0), "runningEpisodeStats");
    }
    std::uintptr_t handle() const { return reinterpret_cast<std::uintptr_t>(h_); }

private:
    irrl_env* h_ = nullptr;
    py::ssize_t n_ = 0;
};

}  // namespace

PYBIND11_MODULE(IRRL_PYMODULE_NAME, m) {
    m.doc() = "FlexibleGymEnv of IRRL's flex_gym (raisim_gym.cpp) over the B200-native CUDA library";
    py::class_<FlexibleGymEnv>(m, "FlexibleGymEnv")
        .def(py::init<std::string, std::string, int, int>(), py::arg("resourceDir"), py::arg("cfg"), py::arg("device") = 0, py::arg("env_offset") = 0)
        .def("init", &FlexibleGymEnv::init)
        .def("getExtraInfoNames", &FlexibleGymEnv::getExtraInfoNames)
        .def("reset", &FlexibleGymEnv::reset)
        .def("observe", &FlexibleGymEnv::observe)
        .def("step", &FlexibleGymEnv::step)
        .def("setSeed", &FlexibleGymEnv::setSeed)
        .def("testStep", &FlexibleGymEnv::testStep)
        .def("close", &FlexibleGymEnv::close)
        .def("isTerminalState", &FlexibleGymEnv::isTerminalState)
        .def("setSimulationTimeStep", &FlexibleGymEnv::setSimulationTimeStep)
        .def("setControlTimeStep", &FlexibleGymEnv::setControlTimeStep)
        .def("getObDim", &FlexibleGymEnv::getObDim)
        .def("getActionDim", &FlexibleGymEnv::getActionDim)
        .def("getExtraInfoDim", &FlexibleGymEnv::getExtraInfoDim)
        .def("getNumOfEnvs", &FlexibleGymEnv::getNumOfEnvs)
        .def("startRecordingVideo", &FlexibleGymEnv::startRecordingVideo)
        .def("stopRecordingVideo", &FlexibleGymEnv::stopRecordingVideo)
        .def("showWindow", &FlexibleGymEnv::showWindow)
        .def("hideWindow", &FlexibleGymEnv::hideWindow)
        .def("curriculumUpdate", &FlexibleGymEnv::curriculumUpdate)
        .def("OriginState", &FlexibleGymEnv::OriginState)
        .def("GetOriginStateDim", &FlexibleGymEnv::GetOriginStateDim)
        .def("ReferenceState", &FlexibleGymEnv::ReferenceState)
        .def("GetJointEffort", &FlexibleGymEnv::GetJointEffort)
        .def("GetGeneralizedForce", &FlexibleGymEnv::GetGeneralizedForce)
        .def("GetInverseMassMatrix", &FlexibleGymEnv::GetInverseMassMatrix)
        .def("GetNonlinear", &FlexibleGymEnv::GetNonlinear)
        .def("SetContactCoefficient", &FlexibleGymEnv::SetContactCoefficient)
        .def("GetSphereInfo", &FlexibleGymEnv::GetSphereInfo)
        .def("lastEpisodeStats", &FlexibleGymEnv::lastEpisodeStats)
        .def("runningEpisodeStats", &FlexibleGymEnv::runningEpisodeStats, py::arg("ep_return"), py::arg("ep_length"), py::arg("clear") = false)
        .def_property_readonly("handle", &FlexibleGymEnv::handle);
}
