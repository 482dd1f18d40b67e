// Low-overhead Python -> C ABI call path for libmolsde_b200 (x86-64 SysV only).
//
// The training step issues ~750 kernel launches per iteration from Python; through ctypes each call costs ~3 us of argument
// marshalling (measured: 21 arguments).  `bind(address, signature)` returns a callable that converts its arguments straight from the
// Python objects according to the DECLARED C types ('i' = any integer / pointer class argument, None -> NULL; 'f' = float) and calls
// the function: ~0.4 us.  Calling convention: on x86-64 SysV integer-class arguments go to rdi, rsi, rdx, rcx, r8, r9 and then to
// the stack in declaration order, float arguments to xmm0..xmm7 -- the two classes are assigned independently, so a function with n
// integer-class and <= 8 float parameters in ANY interleaving can be called as f(int_0 .. int_{n-1}, float_0 .. float_7); unused
// xmm registers are ignored by the callee, 32-bit parameters read the low half of their 64-bit slot.  Every entry point of
// include/molsde_b200.h fits (pointers, int32/int64/uint64, float; status code in eax).  Not used for functions taking ctypes
// structures by reference or returning anything but the int status: those stay on ctypes (moleculesde_b200/_abi.py).
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stddef.h>
#include <string.h>

#if !defined(__x86_64__)
#error "fastcall.c relies on the x86-64 SysV calling convention"
#endif

#define FC_MAX_INTS 28

typedef struct {
    PyObject_HEAD
    vectorcallfunc vectorcall;
    void* fn;
    int nargs;
    char sig[FC_MAX_INTS + 9];
} FcBound;

static PyObject* fc_bound_call(PyObject* self_, PyObject* const* args, size_t nargsf, PyObject* kwnames) {
    FcBound* self = (FcBound*)self_;
    const Py_ssize_t nargs = PyVectorcall_NARGS(nargsf);
    if (kwnames != NULL && PyTuple_GET_SIZE(kwnames) != 0) { PyErr_SetString(PyExc_TypeError, "no keyword arguments"); return NULL; }
    if (nargs != self->nargs) { PyErr_Format(PyExc_TypeError, "expected %d arguments, got %zd", self->nargs, nargs); return NULL; }
    long long iv[FC_MAX_INTS];
    float fv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int ni = 0, nf = 0;
    for (Py_ssize_t a = 0; a < nargs; ++a) {
        PyObject* o = args[a];
        if (self->sig[a] == 'f') {
            double d = PyFloat_Check(o) ? PyFloat_AS_DOUBLE(o) : PyFloat_AsDouble(o);
            if (d == -1.0 && PyErr_Occurred()) return NULL;
            fv[nf++] = (float)d;
        } else if (o == Py_None) {
            iv[ni++] = 0;
        } else {
            unsigned long long v = PyLong_AsUnsignedLongLongMask(o);
            if (v == (unsigned long long)-1 && PyErr_Occurred()) return NULL;
            iv[ni++] = (long long)v;
        }
    }
    void* fn = self->fn;
    int r = 0;
    Py_BEGIN_ALLOW_THREADS
    switch (ni) {
        case 0: r = ((int (*)(float, float, float, float, float, float, float, float))fn)(fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 1: r = ((int (*)(long long, float, float, float, float, float, float, float, float))fn)(iv[0], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 2: r = ((int (*)(long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 3: r = ((int (*)(long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 4: r = ((int (*)(long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 5: r = ((int (*)(long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 6: r = ((int (*)(long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 7: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 8: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 9: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 10: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 11: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 12: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 13: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 14: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 15: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 16: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 17: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], iv[16], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 18: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], iv[16], iv[17], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 19: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], iv[16], iv[17], iv[18], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 20: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], iv[16], iv[17], iv[18], iv[19], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 21: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], iv[16], iv[17], iv[18], iv[19], iv[20], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 22: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], iv[16], iv[17], iv[18], iv[19], iv[20], iv[21], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 23: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], iv[16], iv[17], iv[18], iv[19], iv[20], iv[21], iv[22], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 24: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], iv[16], iv[17], iv[18], iv[19], iv[20], iv[21], iv[22], iv[23], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 25: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], iv[16], iv[17], iv[18], iv[19], iv[20], iv[21], iv[22], iv[23], iv[24], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 26: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], iv[16], iv[17], iv[18], iv[19], iv[20], iv[21], iv[22], iv[23], iv[24], iv[25], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 27: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], iv[16], iv[17], iv[18], iv[19], iv[20], iv[21], iv[22], iv[23], iv[24], iv[25], iv[26], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        case 28: r = ((int (*)(long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, long long, float, float, float, float, float, float, float, float))fn)(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13], iv[14], iv[15], iv[16], iv[17], iv[18], iv[19], iv[20], iv[21], iv[22], iv[23], iv[24], iv[25], iv[26], iv[27], fv[0], fv[1], fv[2], fv[3], fv[4], fv[5], fv[6], fv[7]); break;
        default: r = -1000000; break;
    }
    Py_END_ALLOW_THREADS
    return PyLong_FromLong((long)r);
}

static PyTypeObject FcBoundType = {
    PyVarObject_HEAD_INIT(NULL, 0)
    .tp_name = "_molsde_fastcall.Bound",
    .tp_basicsize = sizeof(FcBound),
    .tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_HAVE_VECTORCALL,
    .tp_vectorcall_offset = offsetof(FcBound, vectorcall),
    .tp_call = PyVectorcall_Call,
};

static PyObject* fc_bind(PyObject* mod, PyObject* args) {
    unsigned long long addr;
    const char* sig;
    if (!PyArg_ParseTuple(args, "Ks", &addr, &sig)) return NULL;
    const size_t n = strlen(sig);
    int ni = 0, nf = 0;
    for (size_t i = 0; i < n; ++i) {
        if (sig[i] == 'f') ++nf; else if (sig[i] == 'i') ++ni; else { PyErr_SetString(PyExc_ValueError, "signature characters: i, f"); return NULL; }
    }
    if (ni > FC_MAX_INTS || nf > 8 || addr == 0) { PyErr_SetString(PyExc_ValueError, "unsupported signature"); return NULL; }
    FcBound* b = PyObject_New(FcBound, &FcBoundType);
    if (!b) return NULL;
    b->vectorcall = fc_bound_call;
    b->fn = (void*)addr;
    b->nargs = (int)n;
    memcpy(b->sig, sig, n + 1);
    return (PyObject*)b;
}

static PyMethodDef fc_methods[] = {{"bind", fc_bind, METH_VARARGS, "bind(address, signature) -> callable returning the int status"}, {NULL, NULL, 0, NULL}};
static struct PyModuleDef fc_module = {PyModuleDef_HEAD_INIT, "_molsde_fastcall", NULL, -1, fc_methods};

PyMODINIT_FUNC PyInit__molsde_fastcall(void) {
    if (PyType_Ready(&FcBoundType) < 0) return NULL;
    return PyModule_Create(&fc_module);
}
