// TEST INFRASTRUCTURE.  Stand-in for 3rdparty/matplotlibcpp.h (matplotlib and its CPython embedding are absent):
// every drawing call is a no-op except text(), which prints what the reference's main loop writes on the figure
// — the ego state and the applied control of every tick — and pause(), which marks the end of a tick.  Also the
// handful of CPython / NumPy C-API names src/utils.cpp's imshow mentions.
#pragma once
#include <cstdio>
#include <map>
#include <string>
#include <vector>

struct PyObject {};
typedef long npy_intp;
#define NPY_FLOAT 11
inline void Py_Initialize() {}
inline int _import_array() { return 0; }
inline int PyRun_SimpleString(const char*) { return 0; }
inline PyObject* PyUnicode_DecodeFSDefault(const char*) { return nullptr; }
inline PyObject* PyImport_Import(PyObject*) { return nullptr; }
inline void Py_DECREF(PyObject*) {}
inline PyObject* PyObject_GetAttrString(PyObject*, const char*) { return nullptr; }
inline int PyCallable_Check(PyObject*) { return 0; }
inline PyObject* PyTuple_New(int) { return nullptr; }
inline int PyTuple_SetItem(PyObject*, int, PyObject*) { return 0; }
inline PyObject* PyArray_SimpleNewFromData(int, npy_intp*, int, void*) { return nullptr; }
inline PyObject* PyObject_CallObject(PyObject*, PyObject*) { return nullptr; }

namespace matplotlibcpp {
typedef std::map<std::string, std::string> Keywords;
inline void cla() {}
inline void show() {}
template <typename A, typename B>
inline bool plot(const A&, const B&, const Keywords& = {}) { return true; }
template <typename A, typename B>
inline bool plot(const A&, const B&, const std::string&) { return true; }
template <typename A, typename B>
inline bool fill(const A&, const B&, const Keywords& = {}) { return true; }
inline void text(double, double, const std::string& s, const Keywords& = {}) { std::printf("TEXT %s\n", s.c_str()); }
inline void xlim(double, double) {}
inline void ylim(double, double) {}
inline void pause(double) { std::printf("TICK\n"); std::fflush(stdout); }
namespace detail {
template <typename T>
inline PyObject* get_array(const std::vector<T>&) { return nullptr; }
}  // namespace detail
}  // namespace matplotlibcpp
