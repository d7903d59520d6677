// ORACLE — TEST INFRASTRUCTURE ONLY.
// Stand-in for 3rdparty/matplotlibcpp.h and the handful of CPython / NumPy C-API names that
// src/utils.cpp's plotting helpers mention, so that file compiles unmodified without Python.
// Nothing here is ever called by the solver path.
#pragma once
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
template <typename A, typename B>
inline bool plot(const A&, const B&, const std::map<std::string, std::string>& = {}) { return true; }
template <typename A, typename B>
inline bool plot(const A&, const B&, const std::string&) { return true; }
namespace detail {
template <typename T>
inline PyObject* get_array(const std::vector<T>&) { return nullptr; }
}  // namespace detail
}  // namespace matplotlibcpp
