/*
 * gbrl_cpp_module.c -- the extension module the reference's loader looks for.
 *
 * gbrl/__init__.py:40-118 globs its package directory for `gbrl_cpp*cpython-3XX*.so`, loads it with
 * importlib and takes `module.GBRL` as `gbrl.GBRL_CPP`.  This file builds exactly such a module
 * (gbrl_b200/lib/gbrl_cpp.<EXT_SUFFIX>): dropped next to the reference's `gbrl/__init__.py` it makes the UNMODIFIED
 * reference package (learners, models, GBRL_SB3 on top) run on the B200 engine.  The class it exports is the host-side
 * mirror of binding.cpp:421-1134 (gbrl_b200/gbrl_cpp.py), which forwards every compute call to the C-ABI of
 * libgbrl_b200.so; there is no CPU implementation behind it.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>

static struct PyModuleDef gbrl_cpp_def = {
    PyModuleDef_HEAD_INIT, "gbrl_cpp",
    "B200-native GBRL engine behind the reference's gbrl_cpp.GBRL surface (see include/gbrl_b200.h)", -1, NULL,
};

PyMODINIT_FUNC PyInit_gbrl_cpp(void) {
    PyObject *impl = PyImport_ImportModule("gbrl_b200.gbrl_cpp");
    if (!impl) return NULL;                      /* ImportError: gbrl_b200 (or libgbrl_b200.so) is not importable */
    PyObject *cls = PyObject_GetAttrString(impl, "GBRL");
    Py_DECREF(impl);
    if (!cls) return NULL;
    PyObject *m = PyModule_Create(&gbrl_cpp_def);
    if (!m) { Py_DECREF(cls); return NULL; }
    if (PyModule_AddObject(m, "GBRL", cls) < 0) { Py_DECREF(cls); Py_DECREF(m); return NULL; }
    PyModule_AddStringConstant(m, "__backend__", "gbrl_b200");
    return m;
}
