// pybind11 module `_kfunca`: the reference's Python surface (src/register.cpp:59-225) re-exposed over the
// C ABI in include/kfunca_b200.h — this file includes nothing else from the library.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kfunca_b200.h"

namespace py = pybind11;

static void ck(int status) {
    if (status != 0) throw std::runtime_error(kf_last_error());
}

// Owning wrapper; copy = new handle on the same impl (ref: register.cpp:89-90, Tensor(self))
struct PyTensor {
    kf_tensor_t h = nullptr;
    PyTensor() = default;
    explicit PyTensor(kf_tensor_t h_) : h(h_) {}
    PyTensor(const PyTensor &o) {
        if (o.h) ck(kf_retain(o.h, &h));
    }
    PyTensor(PyTensor &&o) noexcept : h(o.h) { o.h = nullptr; }
    PyTensor &operator=(PyTensor o) noexcept {
        std::swap(h, o.h);
        return *this;
    }
    ~PyTensor() {
        if (h) kf_release(h);
    }
    kf_tensor_t get() const {
        if (!h) throw std::runtime_error("undefined tensor");
        return h;
    }
};

static int dtype_of_format(const py::array &a) {
    const py::dtype dt = a.dtype();
    const char kind = dt.kind();
    const auto sz = dt.itemsize();
    if (kind == 'b') return KF_BOOL;
    if (kind == 'u' && sz == 1) return KF_BYTE;
    if (kind == 'i' && sz == 1) return KF_CHAR;
    if (kind == 'i' && sz == 2) return KF_SHORT;
    if (kind == 'i' && sz == 4) return KF_INT;
    if (kind == 'i' && sz == 8) return KF_LONG;  // accepts both 'l' and 'q' (the reference only matches 'q')
    if (kind == 'f' && sz == 2) return KF_HALF;
    if (kind == 'f' && sz == 4) return KF_FLOAT;
    if (kind == 'f' && sz == 8) return KF_DOUBLE;
    if (kind == 'V' && sz == 2) {  // ml_dtypes.bfloat16 registers as a 2-byte void-kind dtype named "bfloat16"
        const std::string name = py::str(dt.attr("name"));
        if (name == "bfloat16") return KF_BFLOAT16;
    }
    throw std::runtime_error("Unsupported dtype in from_numpy()");
}

static PyTensor from_numpy(py::array array, int device) {
    const int dtype = dtype_of_format(array);
    // the reference silently assumes C-contiguity (register.cpp:27-35); we make it so
    py::array c = py::array::ensure(array, py::array::c_style | py::array::forcecast);
    if (!c) throw std::runtime_error("from_numpy(): cannot make the array C-contiguous");
    std::vector<int64_t> shape(c.shape(), c.shape() + c.ndim());
    kf_tensor_t h;
    ck(kf_from_host(c.data(), shape.data(), (int)shape.size(), dtype, device, &h));
    return PyTensor(h);
}

static py::object numpy_dtype(int dtype) {
    py::module_ np = py::module_::import("numpy");
    switch (dtype) {
    case KF_BOOL: return np.attr("dtype")("bool");
    case KF_BYTE: return np.attr("dtype")("uint8");
    case KF_CHAR: return np.attr("dtype")("int8");
    case KF_SHORT: return np.attr("dtype")("int16");
    case KF_INT: return np.attr("dtype")("int32");
    case KF_LONG: return np.attr("dtype")("int64");
    case KF_HALF: return np.attr("dtype")("float16");
    case KF_FLOAT: return np.attr("dtype")("float32");
    case KF_DOUBLE: return np.attr("dtype")("float64");
    case KF_BFLOAT16: {
        try {
            return np.attr("dtype")(py::module_::import("ml_dtypes").attr("bfloat16"));
        } catch (py::error_already_set &) {
            throw std::runtime_error("Unsupported dtype in to_numpy(): bfloat16 needs the ml_dtypes package (or call .float() first)");
        }
    }
    default: throw std::runtime_error("Unsupported dtype in to_numpy()");
    }
}

static py::array to_numpy(const PyTensor &t) {
    int contiguous = 0, dtype = 0, ndim = 0;
    ck(kf_is_contiguous(t.get(), &contiguous));
    if (!contiguous) throw std::runtime_error("[enforce fail at to_numpy] `t.is_contiguous()`");
    ck(kf_dtype(t.get(), &dtype));
    int64_t sizes[KF_MAX_DIMS];
    ck(kf_sizes(t.get(), sizes, &ndim));
    std::vector<py::ssize_t> shape(sizes, sizes + ndim);
    py::array out(py::dtype::from_args(numpy_dtype(dtype)), shape);
    ck(kf_to_host(t.get(), out.mutable_data(), (size_t)out.nbytes()));
    return out;
}

static float half_bits_to_float(uint16_t h) {
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, mant = h & 0x3ffu, x;
    if (exp == 0) {
        float f = (float)mant * 5.9604644775390625e-8f;  // mant * 2^-24
        std::memcpy(&x, &f, 4);
        x |= sign;
    } else if (exp == 31) {
        x = sign | 0x7f800000u | (mant << 13);
    } else {
        x = sign | ((exp + 112) << 23) | (mant << 13);
    }
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

static std::vector<int64_t> args_to_dims(const py::args &args) {
    std::vector<int64_t> dims;
    if (args.size() == 1 && (py::isinstance<py::tuple>(args[0]) || py::isinstance<py::list>(args[0]))) {
        for (auto a : args[0]) dims.push_back(a.cast<int64_t>());
    } else {
        for (auto a : args) dims.push_back(a.cast<int64_t>());
    }
    return dims;
}

static PyTensor binary(int op, const PyTensor &a, const PyTensor &b) {
    kf_tensor_t h;
    ck(kf_binary(op, a.get(), b.get(), &h));
    return PyTensor(h);
}
static PyTensor binary_scalar(int op, const PyTensor &a, double s) {
    kf_tensor_t h;
    ck(kf_binary_scalar(op, a.get(), s, &h));
    return PyTensor(h);
}

static PyTensor getitem(const PyTensor &self, py::object key) {
    PyTensor out = self;
    auto apply_slice = [&](const py::slice &s, int dim) {
        int64_t n;
        ck(kf_shape(out.get(), dim, &n));
        size_t start, stop, step, len;
        if (!s.compute((size_t)n, &start, &stop, &step, &len)) throw py::error_already_set();
        if ((py::ssize_t)step <= 0) throw std::runtime_error("slice step must be positive");
        kf_tensor_t h;
        ck(kf_slice(out.get(), dim, (int64_t)start, (int64_t)stop, (int64_t)step, &h));
        out = PyTensor(h);
    };
    auto apply_int = [&](int64_t idx, int dim) {
        kf_tensor_t h;
        ck(kf_select(out.get(), dim, idx, &h));
        out = PyTensor(h);
    };
    if (py::isinstance<py::tuple>(key)) {
        auto t = key.cast<py::tuple>();
        int nd;
        ck(kf_dim(self.get(), &nd));
        if ((int)t.size() > nd) throw std::runtime_error("too many indices for tensor");
        int dim = 0;
        for (auto item : t) {
            if (py::isinstance<py::slice>(item)) {
                apply_slice(item.cast<py::slice>(), dim);
                dim++;
            } else if (py::isinstance<py::int_>(item)) {
                apply_int(item.cast<int64_t>(), dim);
            } else {
                throw std::runtime_error("unsupported index type");
            }
        }
    } else if (py::isinstance<py::slice>(key)) {
        apply_slice(key.cast<py::slice>(), 0);
    } else {
        apply_int(key.cast<int64_t>(), 0);
    }
    return out;
}

PYBIND11_MODULE(_kfunca, m) {
    m.doc() = "kfunca_b200: B200-native implementation of the kfunca tensor-operator API";
    m.def("device_info", []() {
        std::vector<char> buf(4096);
        ck(kf_device_info(buf.data(), buf.size()));
        py::print(std::string(buf.data()));
    });
    m.def("memstat", []() {
        std::vector<char> buf(1 << 16);
        ck(kf_memstat(buf.data(), buf.size()));
        py::print(std::string(buf.data()));
    });
    m.def("mem_stats", []() {
        int64_t a, b, c;
        ck(kf_mem_stats(&a, &b, &c));
        return py::make_tuple(a, b, c);
    });
    m.def("empty_cache", []() { ck(kf_empty_cache()); });
    m.def("synchronize", []() { ck(kf_synchronize()); });
    m.def("set_device", [](int d) { ck(kf_set_device(d)); });
    m.def("get_device", []() { int d; ck(kf_get_device(&d)); return d; });
    m.def("device_count", []() { int n; ck(kf_device_count(&n)); return n; });
    m.def("launch_count", []() { int64_t n; ck(kf_launch_count(&n)); return n; });
    // set_leaf_grad_hook(fn | None): fn(leaf, grad) is called inside backward() as each leaf gradient is enqueued (see the header)
    m.def("set_leaf_grad_hook", [](py::object fn) {
        static py::object *held = new py::object();  // leaked on purpose: must outlive interpreter finalisation order
        if (fn.is_none()) {
            ck(kf_set_leaf_grad_hook(nullptr, nullptr));
            *held = py::none();
            return;
        }
        *held = fn;
        ck(kf_set_leaf_grad_hook(
            [](kf_tensor_t leaf, kf_tensor_t grad, void *ctx) {
                py::gil_scoped_acquire gil;
                kf_tensor_t l2 = nullptr, g2 = nullptr;  // owning handles for Python
                if (kf_retain(leaf, &l2) != 0 || kf_retain(grad, &g2) != 0) return;
                try {
                    (*static_cast<py::object *>(ctx))(PyTensor(l2), PyTensor(g2));
                } catch (py::error_already_set &e) {  // a Python error must not unwind through the C ABI
                    e.restore();
                    PyErr_Print();
                }
            },
            held));
    });
    m.def("stream", []() { void *s; ck(kf_stream(&s)); return (uintptr_t)s; });

    py::enum_<kf_dtype_t>(m, "dtype")
        .value("byte", KF_BYTE)
        .value("char", KF_CHAR)
        .value("short", KF_SHORT)
        .value("int", KF_INT)
        .value("long", KF_LONG)
        .value("half", KF_HALF)
        .value("bfloat16", KF_BFLOAT16)
        .value("float", KF_FLOAT)
        .value("double", KF_DOUBLE)
        .value("bool", KF_BOOL)
        .export_values();

    m.def("empty", [](std::vector<int64_t> shape, kf_dtype_t dtype, int device) {
        kf_tensor_t h;
        ck(kf_empty(shape.data(), (int)shape.size(), dtype, device, &h));
        return PyTensor(h);
    });
    m.def("zeros", [](std::vector<int64_t> shape, kf_dtype_t dtype, int device) {
        kf_tensor_t h;
        ck(kf_zeros(shape.data(), (int)shape.size(), dtype, device, &h));
        return PyTensor(h);
    });
    m.def("empty_like", [](const PyTensor &t) {
        kf_tensor_t h;
        ck(kf_empty_like(t.get(), &h));
        return PyTensor(h);
    });
    m.def("from_numpy", &from_numpy);
    m.def("to_numpy", &to_numpy);
    m.def("causal_attention", [](const PyTensor &q, const PyTensor &k, const PyTensor &v) {
        kf_tensor_t h;
        ck(kf_causal_attention(q.get(), k.get(), v.get(), &h));
        return PyTensor(h);
    });
    m.def("causal_attention_fwd", [](const PyTensor &q, const PyTensor &k, const PyTensor &v) {
        kf_tensor_t o, l;
        ck(kf_causal_attention_fwd(q.get(), k.get(), v.get(), &o, &l));
        return py::make_tuple(PyTensor(o), PyTensor(l));
    });
    m.def("causal_attention_bwd", [](const PyTensor &dout, const PyTensor &q, const PyTensor &k, const PyTensor &v, const PyTensor &out,
                                     const PyTensor &lse) {
        kf_tensor_t a, b, c;
        ck(kf_causal_attention_bwd(dout.get(), q.get(), k.get(), v.get(), out.get(), lse.get(), &a, &b, &c));
        return py::make_tuple(PyTensor(a), PyTensor(b), PyTensor(c));
    });
    m.def("gemm", [](const PyTensor &a, const PyTensor &b, float alpha, float beta) {
        kf_tensor_t h;
        ck(kf_gemm(a.get(), b.get(), alpha, beta, &h));
        return PyTensor(h);
    });
    m.def("gemm_out", [](PyTensor &out, const PyTensor &a, const PyTensor &b, float alpha, float beta) {
        ck(kf_gemm_out(out.get(), a.get(), b.get(), alpha, beta));
    });
    m.def("matmul", [](const PyTensor &a, bool ta, const PyTensor &b, bool tb, float alpha) {
        kf_tensor_t h;
        ck(kf_matmul(a.get(), ta, b.get(), tb, alpha, &h));
        return PyTensor(h);
    }, py::arg("a"), py::arg("trans_a"), py::arg("b"), py::arg("trans_b"), py::arg("alpha") = 1.0f);
    m.def("cat", [](std::vector<PyTensor> ts, int64_t dim) {
        std::vector<kf_tensor_t> hs;
        for (auto &t : ts) hs.push_back(t.get());
        kf_tensor_t h;
        ck(kf_cat(hs.data(), (int)hs.size(), dim, &h));
        return PyTensor(h);
    });
    m.def("sqrt", [](const PyTensor &a) { kf_tensor_t h; ck(kf_unary(KF_UOP_SQRT, a.get(), &h)); return PyTensor(h); });
    m.def("layer_norm", [](const PyTensor &x, const PyTensor &gain, double eps) {
        kf_tensor_t h;
        ck(kf_layer_norm(x.get(), gain.get(), eps, &h));
        return PyTensor(h);
    });
    // data-parallel layer (csrc/dist.cpp)
    m.def("dist_unique_id", []() {
        char id[128];
        ck(kf_dist_unique_id(id));
        return py::bytes(id, 128);
    });
    m.def("dist_init", [](py::bytes id, int rank, int world) {
        std::string s = id;
        if (s.size() != 128) throw std::runtime_error("dist_init: the id must be 128 bytes");
        ck(kf_dist_init(s.data(), rank, world));
    });
    m.def("dist_finalize", []() { ck(kf_dist_finalize()); });
    m.def("dist_info", []() {
        int i, r, w, v;
        ck(kf_dist_info(&i, &r, &w, &v));
        return py::make_tuple(i != 0, r, w, v);
    });
    m.def("dist_all_reduce", [](PyTensor &t, int op) { ck(kf_dist_all_reduce(t.get(), op)); return t; });
    m.def("dist_overlap_begin", [](std::vector<PyTensor> params) {
        std::vector<kf_tensor_t> hs;
        for (auto &t : params) hs.push_back(t.get());
        ck(kf_dist_overlap_begin(hs.data(), (int)hs.size()));
    });
    m.def("dist_overlap_end", []() { int64_t n; ck(kf_dist_overlap_end(&n)); return n; });
    m.def("rms_norm", [](const PyTensor &x, const PyTensor &gain, double eps) {
        kf_tensor_t h;
        ck(kf_rms_norm(x.get(), gain.get(), eps, &h));
        return PyTensor(h);
    });
    m.def("gemm_residual", [](const PyTensor &a, const PyTensor &b, const PyTensor &r, float alpha) {
        kf_tensor_t h;
        ck(kf_gemm_residual(a.get(), b.get(), r.get(), alpha, &h));
        return PyTensor(h);
    });
    m.def("qkv_linear", [](const PyTensor &x, const PyTensor &w, py::object bias) {  // bias: tensor or None
        kf_tensor_t h;
        ck(kf_qkv_linear(x.get(), w.get(), bias.is_none() ? nullptr : bias.cast<const PyTensor &>().get(), &h));
        return PyTensor(h);
    }, py::arg("x"), py::arg("w"), py::arg("bias") = py::none());
    m.def("gemm_glu", [](const PyTensor &a, const PyTensor &b1, const PyTensor &b3) {
        kf_tensor_t h;
        ck(kf_gemm_glu(a.get(), b1.get(), b3.get(), &h));
        return PyTensor(h);
    });
    m.def("qkv_attention", [](const PyTensor &qkv, int64_t heads) {
        kf_tensor_t h;
        ck(kf_qkv_attention(qkv.get(), heads, &h));
        return PyTensor(h);
    });
    m.def("embedding", [](const PyTensor &w, const PyTensor &idx) {
        kf_tensor_t h;
        ck(kf_embedding(w.get(), idx.get(), &h));
        return PyTensor(h);
    });
    m.def("rsqrt", [](const PyTensor &a) { kf_tensor_t h; ck(kf_unary(KF_UOP_RSQRT, a.get(), &h)); return PyTensor(h); });
    m.def("neg", [](const PyTensor &a) { kf_tensor_t h; ck(kf_unary(KF_UOP_NEG, a.get(), &h)); return PyTensor(h); });
    m.def("promote_types", [](kf_dtype_t a, kf_dtype_t b) { int o; ck(kf_promote_types(a, b, &o)); return (kf_dtype_t)o; });
    m.def("debug_plan_binary", [](const PyTensor &a, const PyTensor &b) {
        int ndim, common;
        int64_t shape[KF_MAX_DIMS], strides[3 * KF_MAX_DIMS];
        ck(kf_debug_plan_binary(a.get(), b.get(), &ndim, shape, strides, &common));
        std::vector<int64_t> sh(shape, shape + ndim);
        std::vector<std::vector<int64_t>> st(3);
        for (int i = 0; i < 3; ++i) st[i].assign(strides + i * KF_MAX_DIMS, strides + i * KF_MAX_DIMS + ndim);
        return py::make_tuple(sh, st, (kf_dtype_t)common);
    });
    m.def("debug_pool_fences", [](std::vector<int> kind, std::vector<int64_t> arg, std::vector<int64_t> arg2) {
        std::vector<int64_t> log(kind.size() * 4 + 4);
        int n = 0;
        ck(kf_debug_pool_fences(kind.data(), arg.data(), arg2.data(), (int)kind.size(), log.data(), (int)log.size(), &n));
        log.resize(std::min<size_t>((size_t)n, log.size()));
        return py::make_tuple(log, n);
    });
    m.def("numa_info", []() {
        int node = -1;
        char buf[1024] = {0};
        ck(kf_numa_info(&node, buf, sizeof(buf)));
        return py::make_tuple(node, std::string(buf));
    });
    m.def("record_stream", [](const PyTensor &t, uintptr_t stream) { ck(kf_record_stream(t.get(), reinterpret_cast<void *>(stream))); });
    m.def("debug_pool_trace", [](std::vector<int64_t> ops) {
        std::vector<int64_t> offs(ops.size());
        int64_t stats[3];
        ck(kf_debug_pool_trace(ops.data(), (int)ops.size(), offs.data(), stats));
        return py::make_tuple(offs, std::vector<int64_t>(stats, stats + 3));
    });

    py::class_<PyTensor>(m, "tensor")
        .def("__copy__", [](const PyTensor &self) { return PyTensor(self); })
        .def("__deepcopy__", [](const PyTensor &self, py::dict) { return PyTensor(self); })
        .def("__repr__", [](const PyTensor &self) {
            std::vector<char> buf(1 << 16);
            ck(kf_to_string(self.h, buf.data(), buf.size()));
            return std::string(buf.data());
        })
        .def("defined", [](const PyTensor &self) { int d = 0; if (self.h) ck(kf_defined(self.h, &d)); return d != 0; })
        .def("numpy", [](const PyTensor &self) { return to_numpy(self); })
        .def("numel", [](const PyTensor &self) { int64_t n; ck(kf_numel(self.get(), &n)); return n; })
        .def("dim", [](const PyTensor &self) { int n; ck(kf_dim(self.get(), &n)); return n; })
        .def("device", [](const PyTensor &self) { int n; ck(kf_device(self.get(), &n)); return n; })
        .def("shape", [](const PyTensor &self, int64_t d) { int64_t n; ck(kf_shape(self.get(), (int)d, &n)); return n; })
        .def("sizes", [](const PyTensor &self) {
            int64_t s[KF_MAX_DIMS]; int nd; ck(kf_sizes(self.get(), s, &nd));
            return std::vector<int64_t>(s, s + nd);
        })
        .def("strides", [](const PyTensor &self) {
            int64_t s[KF_MAX_DIMS]; int nd; ck(kf_strides(self.get(), s, &nd));
            return std::vector<int64_t>(s, s + nd);
        })
        .def("storage_offset", [](const PyTensor &self) { int64_t n; ck(kf_storage_offset(self.get(), &n)); return n; })
        .def("is_contiguous", [](const PyTensor &self) { int n; ck(kf_is_contiguous(self.get(), &n)); return n != 0; })
        .def("dtype", [](const PyTensor &self) { int n; ck(kf_dtype(self.get(), &n)); return (kf_dtype_t)n; })
        .def("item", [](const PyTensor &self, std::vector<int64_t> indices) -> py::object {
            int dtype;
            ck(kf_dtype(self.get(), &dtype));
            alignas(8) unsigned char raw[8];
            ck(kf_item(self.get(), indices.data(), (int)indices.size(), raw));
            switch (dtype) {
            case KF_BOOL: return py::cast(raw[0] != 0);
            case KF_BYTE: return py::cast(*reinterpret_cast<uint8_t *>(raw));
            case KF_CHAR: return py::cast(*reinterpret_cast<int8_t *>(raw));
            case KF_SHORT: return py::cast(*reinterpret_cast<int16_t *>(raw));
            case KF_INT: return py::cast(*reinterpret_cast<int32_t *>(raw));
            case KF_LONG: return py::cast(*reinterpret_cast<int64_t *>(raw));
            case KF_FLOAT: return py::cast(*reinterpret_cast<float *>(raw));
            case KF_DOUBLE: return py::cast(*reinterpret_cast<double *>(raw));
            case KF_HALF: return py::cast(half_bits_to_float(*reinterpret_cast<uint16_t *>(raw)));
            case KF_BFLOAT16: {
                uint32_t x = (uint32_t)(*reinterpret_cast<uint16_t *>(raw)) << 16;
                float f;
                std::memcpy(&f, &x, 4);
                return py::cast(f);
            }
            default: return py::none();
            }
        })
        .def("fill_", [](PyTensor &self, double v) { ck(kf_fill_(self.get(), v)); return self; })
        .def("data_ptr", [](const PyTensor &self) { void *p; ck(kf_data_ptr(self.get(), &p)); return (uintptr_t)p; })
        .def("storage_ref_count", [](const PyTensor &self) { int64_t n; ck(kf_storage_ref_count(self.get(), &n)); return n; })
        .def("impl_ref_count", [](const PyTensor &self) { int64_t n; ck(kf_impl_ref_count(self.get(), &n)); return n; })
        .def("contiguous", [](const PyTensor &self) { kf_tensor_t h; ck(kf_contiguous(self.get(), &h)); return PyTensor(h); })
        .def("clone", [](const PyTensor &self) { kf_tensor_t h; ck(kf_clone(self.get(), &h)); return PyTensor(h); })
        .def("copy_", [](PyTensor &self, const PyTensor &src) { ck(kf_copy_(self.get(), src.get())); return self; })
        .def("permute", [](const PyTensor &self, py::args args) {
            auto dims = args_to_dims(args);
            kf_tensor_t h;
            ck(kf_permute(self.get(), dims.data(), (int)dims.size(), &h));
            return PyTensor(h);
        })
        .def("view", [](const PyTensor &self, py::args args) {
            auto dims = args_to_dims(args);
            kf_tensor_t h;
            ck(kf_view(self.get(), dims.data(), (int)dims.size(), &h));
            return PyTensor(h);
        })
        .def("split", [](const PyTensor &self, std::vector<int64_t> sizes, int64_t dim) {
            std::vector<kf_tensor_t> hs(sizes.size());
            ck(kf_split(self.get(), sizes.data(), (int)sizes.size(), dim, hs.data()));
            std::vector<PyTensor> out;
            for (auto h : hs) out.emplace_back(h);
            return out;
        })
        .def("sort", [](const PyTensor &self, int64_t dim, bool descending) {
            kf_tensor_t v, i;
            ck(kf_sort(self.get(), dim, descending, &v, &i));
            return py::make_tuple(PyTensor(v), PyTensor(i));
        })
        .def("topk", [](const PyTensor &self, int64_t k, int64_t dim, bool largest) {
            kf_tensor_t v, i;
            ck(kf_topk(self.get(), k, dim, largest, &v, &i));
            return py::make_tuple(PyTensor(v), PyTensor(i));
        })
        .def("__getitem__", &getitem)
        .def("__add__", [](const PyTensor &a, const PyTensor &b) { return binary(KF_OP_ADD, a, b); })
        .def("__add__", [](const PyTensor &a, double s) { return binary_scalar(KF_OP_ADD, a, s); })
        .def("__iadd__", [](PyTensor &a, const PyTensor &b) { ck(kf_binary_(KF_OP_ADD, a.get(), b.get())); return a; })
        .def("__iadd__", [](PyTensor &a, double s) { ck(kf_binary_scalar_(KF_OP_ADD, a.get(), s)); return a; })
        .def("__sub__", [](const PyTensor &a, const PyTensor &b) { return binary(KF_OP_SUB, a, b); })
        .def("__sub__", [](const PyTensor &a, double s) { return binary_scalar(KF_OP_SUB, a, s); })
        .def("__isub__", [](PyTensor &a, const PyTensor &b) { ck(kf_binary_(KF_OP_SUB, a.get(), b.get())); return a; })
        .def("__isub__", [](PyTensor &a, double s) { ck(kf_binary_scalar_(KF_OP_SUB, a.get(), s)); return a; })
        .def("__mul__", [](const PyTensor &a, const PyTensor &b) { return binary(KF_OP_MUL, a, b); })
        .def("__mul__", [](const PyTensor &a, double s) { return binary_scalar(KF_OP_MUL, a, s); })
        .def("__imul__", [](PyTensor &a, const PyTensor &b) { ck(kf_binary_(KF_OP_MUL, a.get(), b.get())); return a; })
        .def("__imul__", [](PyTensor &a, double s) { ck(kf_binary_scalar_(KF_OP_MUL, a.get(), s)); return a; })
        .def("__truediv__", [](const PyTensor &a, const PyTensor &b) { return binary(KF_OP_DIV, a, b); })
        .def("__truediv__", [](const PyTensor &a, double s) { return binary_scalar(KF_OP_DIV, a, s); })
        .def("__itruediv__", [](PyTensor &a, const PyTensor &b) { ck(kf_binary_(KF_OP_DIV, a.get(), b.get())); return a; })
        .def("__itruediv__", [](PyTensor &a, double s) { ck(kf_binary_scalar_(KF_OP_DIV, a.get(), s)); return a; })
        .def("sum", [](const PyTensor &self, int64_t dim) { kf_tensor_t h; ck(kf_sum(self.get(), dim, &h)); return PyTensor(h); })
        .def("mean", [](const PyTensor &self, int64_t dim) { kf_tensor_t h; ck(kf_mean(self.get(), dim, &h)); return PyTensor(h); })
        .def("mean_var", [](const PyTensor &self, int64_t dim, bool take_sqrt) {
            kf_tensor_t a, b;
            ck(kf_mean_var(self.get(), dim, take_sqrt, &a, &b));
            return py::make_tuple(PyTensor(a), PyTensor(b));
        })
        .def("norm_stat", [](const PyTensor &self, int64_t dim) {
            kf_tensor_t a, b;
            ck(kf_norm_stat(self.get(), dim, &a, &b));
            return py::make_tuple(PyTensor(a), PyTensor(b));
        })
        .def("index_put_", [](PyTensor &self, std::vector<PyTensor> indices, const PyTensor &values) {
            std::vector<kf_tensor_t> hs;
            for (auto &t : indices) hs.push_back(t.get());
            ck(kf_index_put_(self.get(), hs.data(), (int)hs.size(), values.get()));
            return self;
        })
        .def("random_uniform_", [](PyTensor &self, uint64_t seed, double lo, double hi) {
            ck(kf_random_uniform_(self.get(), seed, lo, hi));
            return self;
        })
        .def("half", [](const PyTensor &self) { kf_tensor_t h; ck(kf_convert(self.get(), KF_HALF, &h)); return PyTensor(h); })
        .def("bfloat16", [](const PyTensor &self) { kf_tensor_t h; ck(kf_convert(self.get(), KF_BFLOAT16, &h)); return PyTensor(h); })
        .def("float", [](const PyTensor &self) { kf_tensor_t h; ck(kf_convert(self.get(), KF_FLOAT, &h)); return PyTensor(h); })
        .def("double", [](const PyTensor &self) { kf_tensor_t h; ck(kf_convert(self.get(), KF_DOUBLE, &h)); return PyTensor(h); })
        .def("to", [](const PyTensor &self, kf_dtype_t dt) { kf_tensor_t h; ck(kf_convert(self.get(), dt, &h)); return PyTensor(h); })
        .def("requires_grad", [](const PyTensor &self) { int f; ck(kf_requires_grad(self.get(), &f)); return f != 0; })
        .def("set_requires_grad", [](PyTensor &self, bool f) { ck(kf_set_requires_grad(self.get(), f)); })
        .def("backward", [](PyTensor &self, const PyTensor &g) { ck(kf_backward(self.get(), g.get())); })
        .def("zero_grad", [](PyTensor &self) { ck(kf_zero_grad(self.get())); })
        .def("grad", [](const PyTensor &self) {
            kf_tensor_t h = nullptr;
            ck(kf_grad(self.get(), &h));
            return PyTensor(h);
        });
}
