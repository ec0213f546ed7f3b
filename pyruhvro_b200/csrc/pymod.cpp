// CPython extension `pyruhvro_b200._native`: the part of the reference's PyO3 layer that has to
// be native — walking a Python list[bytes] and releasing the GIL around the decode.
//
// Mirrors src/lib.rs of the reference:
//   extract_bytes_list  (:29-33)  elements must be `bytes`; references are held for the call
//   py.detach(...)      (:64-69,82-87)  the GIL is released around packing + decode
//   to_py_err           (:25-27)  failures surface as ValueError(str)
// Packing into one contiguous buffer + offsets is what the reference itself does with
// BinaryArray::from_vec (ruhvro/src/deserialize.rs:90); here it lands in pinned host memory so
// the H2D copy runs at PCIe speed.
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "ruhvro_b200.h"

namespace {

struct Pinned {
    void* p = nullptr;
    explicit Pinned(size_t n) : p(rv_host_alloc(n)) {}
    ~Pinned() { rv_host_free(p); }
};

// decode_list(schema_handle: int, records: list[bytes], num_chunks: int) -> result handle (int)
PyObject* decode_list(PyObject*, PyObject* args) {
    unsigned long long schema_addr = 0;
    PyObject* list = nullptr;
    long long num_chunks = 1;
    if (!PyArg_ParseTuple(args, "KO!L", &schema_addr, &PyList_Type, &list, &num_chunks)) return nullptr;
    const rv_schema* schema = reinterpret_cast<const rv_schema*>(static_cast<uintptr_t>(schema_addr));
    const Py_ssize_t n = PyList_GET_SIZE(list);

    // Pass 1 (GIL held): type-check, take references, record pointers and sizes.
    std::vector<PyObject*> held(static_cast<size_t>(n));
    std::vector<const char*> ptrs(static_cast<size_t>(n));
    std::vector<int64_t> offsets(static_cast<size_t>(n) + 1);
    offsets[0] = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* item = PyList_GET_ITEM(list, i);
        if (!PyBytes_Check(item)) {
            for (Py_ssize_t j = 0; j < i; ++j) Py_DECREF(held[static_cast<size_t>(j)]);
            PyErr_Format(PyExc_TypeError, "argument 'list': element %zd is '%s', expected 'bytes'", i, Py_TYPE(item)->tp_name);
            return nullptr;
        }
        Py_INCREF(item);
        held[static_cast<size_t>(i)] = item;
        ptrs[static_cast<size_t>(i)] = PyBytes_AS_STRING(item);
        offsets[static_cast<size_t>(i) + 1] = offsets[static_cast<size_t>(i)] + static_cast<int64_t>(PyBytes_GET_SIZE(item));
    }
    const int64_t total = offsets[static_cast<size_t>(n)];

    rv_result* result = nullptr;
    rv_status st = RV_OK;
    bool oom = false;
    Py_BEGIN_ALLOW_THREADS;
    {
        Pinned data(static_cast<size_t>(total) + 64), offs((static_cast<size_t>(n) + 1) * 8);
        if (!data.p || !offs.p) {
            oom = true;
        } else {
            std::memcpy(offs.p, offsets.data(), (static_cast<size_t>(n) + 1) * 8);
            // Pass 2 (GIL released): gather the payloads, in parallel for large inputs.
            unsigned workers = total > (int64_t(8) << 20) ? std::min(16u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
            auto copy_range = [&](Py_ssize_t a, Py_ssize_t b) {
                char* dst = static_cast<char*>(data.p);
                for (Py_ssize_t i = a; i < b; ++i)
                    std::memcpy(dst + offsets[static_cast<size_t>(i)], ptrs[static_cast<size_t>(i)],
                                static_cast<size_t>(offsets[static_cast<size_t>(i) + 1] - offsets[static_cast<size_t>(i)]));
            };
            if (workers <= 1) {
                copy_range(0, n);
            } else {
                std::vector<std::thread> th;
                for (unsigned w = 0; w < workers; ++w)
                    th.emplace_back(copy_range, n * Py_ssize_t(w) / Py_ssize_t(workers), n * Py_ssize_t(w + 1) / Py_ssize_t(workers));
                for (auto& t : th) t.join();
            }
            st = rv_decode_host(schema, static_cast<const uint8_t*>(data.p), static_cast<const int64_t*>(offs.p), n, num_chunks, &result);
        }
    }
    Py_END_ALLOW_THREADS;
    for (PyObject* o : held) Py_DECREF(o);
    if (oom) {
        PyErr_SetString(PyExc_ValueError, "pinned host allocation failed (is a CUDA device present?)");
        return nullptr;
    }
    if (st != RV_OK) {
        PyErr_SetString(PyExc_ValueError, rv_last_error());
        return nullptr;
    }
    return PyLong_FromUnsignedLongLong(static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(result)));
}

PyMethodDef methods[] = {
    {"decode_list", decode_list, METH_VARARGS, "decode_list(schema_handle, records: list[bytes], num_chunks) -> result handle"},
    {nullptr, nullptr, 0, nullptr},
};

PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_native", "list[bytes] packing + GIL release for pyruhvro_b200", -1, methods, nullptr, nullptr, nullptr, nullptr};

}  // namespace

PyMODINIT_FUNC PyInit__native(void) { return PyModule_Create(&moddef); }
