/*
 * Minimal stand-in for the C API of `cvxopt.base` — TEST INFRASTRUCTURE ONLY.
 *
 * Why: the only native code of the reference on the Newton-system path is its C extension
 * `smcp.misc` (/root/reference/src/C/misc.c: SCMcolumn2 620-663, Av_to_spmatrix 475-521,
 * scal_diag 542-557, nzcolumns 682-730, matperm 750-773, phase1_sdp 1004-1054, ind2sub /
 * sub2ind 387-447).  It compiles from its single source file against the vendored header
 * /root/reference/src/C/cvxopt.h, but importing it needs `cvxopt.base._C_API`
 * (cvxopt.h:94-113) and cvxopt is neither vendored nor installable here.  The extension
 * touches cvxopt objects only through the two struct layouts and the eight capsule entries
 * that header declares, so this file provides exactly that ABI — a `matrix` type, an
 * `spmatrix` type, and a capsule "base_API" with {Matrix_New, Matrix_NewFromMatrix,
 * Matrix_NewFromList, Matrix_Check, SpMatrix_New, SpMatrix_NewFromSpMatrix,
 * SpMatrix_NewFromIJV, SpMatrix_Check} — plus byte-level constructors/accessors so that
 * oracle/ref.py can move NumPy arrays in and out.  With it, oracle/Makefile builds the
 * UNMODIFIED reference source into oracle/_ref/misc.so and the tests compare our
 * restatements (smcp_b200/misc.py, the CUDA Schur kernel) with the reference's own code.
 *
 * Nothing here is derived from cvxopt's implementation: the layouts are the ones written
 * in the reference's header, everything else is new and deliberately tiny (no arithmetic,
 * no slicing, no printing).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdlib.h>
#include <string.h>

#define INT 0
#define DOUBLE 1
#define COMPLEX 2
typedef Py_ssize_t int_t;

/* layouts: /root/reference/src/C/cvxopt.h:48-69 */
typedef struct {
    PyObject_HEAD
    void *buffer;
    int nrows, ncols;
    int id;
    int_t shape[2];
    int_t strides[2];
    int_t ob_exports;
} matrix;

typedef struct {
    void *values;
    int_t *colptr;
    int_t *rowind;
    int_t nrows, ncols;
    int id;
} ccs;

typedef struct {
    PyObject_HEAD
    ccs *obj;
} spmatrix;

static PyTypeObject matrix_tp, spmatrix_tp;   /* defined below */

static size_t elem_size(int id) { return id == INT ? sizeof(int_t) : (id == DOUBLE ? sizeof(double) : 2 * sizeof(double)); }

/* ---- matrix ---------------------------------------------------------------------------- */
static matrix *Matrix_New(int nrows, int ncols, int id) {
    if (nrows < 0 || ncols < 0 || id < INT || id > COMPLEX) {
        PyErr_SetString(PyExc_TypeError, "illegal matrix dimensions or type");
        return NULL;
    }
    matrix *a = (matrix *)matrix_tp.tp_alloc(&matrix_tp, 0);
    if (!a) return NULL;
    size_t n = (size_t)nrows * (size_t)ncols;
    a->buffer = calloc(n ? n : 1, elem_size(id));
    if (!a->buffer) { Py_DECREF(a); return (matrix *)PyErr_NoMemory(); }
    a->nrows = nrows; a->ncols = ncols; a->id = id;
    a->shape[0] = nrows; a->shape[1] = ncols;
    a->strides[0] = (int_t)elem_size(id); a->strides[1] = (int_t)(elem_size(id) * nrows);
    a->ob_exports = 0;
    return a;
}

static matrix *Matrix_NewFromMatrix(matrix *src, int id) {
    if (id != src->id) { PyErr_SetString(PyExc_TypeError, "shim: type conversion not supported"); return NULL; }
    matrix *a = Matrix_New(src->nrows, src->ncols, id);
    if (a) memcpy(a->buffer, src->buffer, (size_t)src->nrows * src->ncols * elem_size(id));
    return a;
}

static matrix *Matrix_NewFromList(PyObject *list, int id) {
    (void)list; (void)id;
    PyErr_SetString(PyExc_NotImplementedError, "shim: Matrix_NewFromList");
    return NULL;
}

static int Matrix_Check_func(void *o) { return PyObject_TypeCheck((PyObject *)o, &matrix_tp); }

static void matrix_dealloc(matrix *a) {
    free(a->buffer);
    Py_TYPE(a)->tp_free((PyObject *)a);
}

static PyObject *matrix_tobytes(matrix *a, PyObject *unused) {
    (void)unused;
    return PyBytes_FromStringAndSize((const char *)a->buffer, (Py_ssize_t)((size_t)a->nrows * a->ncols * elem_size(a->id)));
}
static PyObject *matrix_get_size(matrix *a, void *c) { (void)c; return Py_BuildValue("ii", a->nrows, a->ncols); }
static PyObject *matrix_get_id(matrix *a, void *c) { (void)c; return PyLong_FromLong(a->id); }

static PyMethodDef matrix_methods[] = {
    {"tobytes", (PyCFunction)matrix_tobytes, METH_NOARGS, "column-major contents"},
    {NULL, NULL, 0, NULL}};
static PyGetSetDef matrix_getset[] = {
    {"size", (getter)matrix_get_size, NULL, "(nrows, ncols)", NULL},
    {"id", (getter)matrix_get_id, NULL, "0 int, 1 double", NULL},
    {NULL, NULL, NULL, NULL, NULL}};

/* base.matrix_from(bytes-like, nrows, ncols, id): column-major contents copied in */
static PyObject *py_matrix_from(PyObject *self, PyObject *args) {
    (void)self;
    Py_buffer view;
    int nrows, ncols, id;
    if (!PyArg_ParseTuple(args, "y*iii", &view, &nrows, &ncols, &id)) return NULL;
    matrix *a = Matrix_New(nrows, ncols, id);
    if (a) {
        size_t want = (size_t)nrows * ncols * elem_size(id);
        if ((size_t)view.len != want) {
            Py_DECREF(a);
            a = NULL;
            PyErr_SetString(PyExc_ValueError, "shim: buffer length does not match the shape");
        } else memcpy(a->buffer, view.buf, want);
    }
    PyBuffer_Release(&view);
    return (PyObject *)a;
}

/* ---- spmatrix -------------------------------------------------------------------------- */
static spmatrix *SpMatrix_New(int_t nrows, int_t ncols, int_t nnz, int id) {
    spmatrix *a = (spmatrix *)spmatrix_tp.tp_alloc(&spmatrix_tp, 0);
    if (!a) return NULL;
    ccs *o = (ccs *)malloc(sizeof(ccs));
    if (!o) { Py_DECREF(a); return (spmatrix *)PyErr_NoMemory(); }
    o->nrows = nrows; o->ncols = ncols; o->id = id;
    o->values = calloc(nnz ? nnz : 1, elem_size(id));
    o->colptr = (int_t *)calloc((size_t)ncols + 1, sizeof(int_t));
    o->rowind = (int_t *)calloc(nnz ? nnz : 1, sizeof(int_t));
    a->obj = o;
    if (!o->values || !o->colptr || !o->rowind) { Py_DECREF(a); return (spmatrix *)PyErr_NoMemory(); }
    return a;
}

static spmatrix *SpMatrix_NewFromSpMatrix(spmatrix *src, int id) {
    ccs *s = src->obj;
    int_t nnz = s->colptr[s->ncols];
    if (id != s->id) { PyErr_SetString(PyExc_TypeError, "shim: type conversion not supported"); return NULL; }
    spmatrix *a = SpMatrix_New(s->nrows, s->ncols, nnz, id);
    if (!a) return NULL;
    memcpy(a->obj->values, s->values, (size_t)nnz * elem_size(id));
    memcpy(a->obj->colptr, s->colptr, ((size_t)s->ncols + 1) * sizeof(int_t));
    memcpy(a->obj->rowind, s->rowind, (size_t)nnz * sizeof(int_t));
    return a;
}

static spmatrix *SpMatrix_NewFromIJV(matrix *I, matrix *J, matrix *V, int_t m, int_t n, int id) {
    (void)I; (void)J; (void)V; (void)m; (void)n; (void)id;
    PyErr_SetString(PyExc_NotImplementedError, "shim: SpMatrix_NewFromIJV");
    return NULL;
}

static int SpMatrix_Check_func(void *o) { return PyObject_TypeCheck((PyObject *)o, &spmatrix_tp); }

static void spmatrix_dealloc(spmatrix *a) {
    if (a->obj) {
        free(a->obj->values);
        free(a->obj->colptr);
        free(a->obj->rowind);
        free(a->obj);
    }
    Py_TYPE(a)->tp_free((PyObject *)a);
}

static PyObject *sp_values(spmatrix *a, PyObject *u) {
    (void)u;
    ccs *o = a->obj;
    return PyBytes_FromStringAndSize((const char *)o->values, (Py_ssize_t)((size_t)o->colptr[o->ncols] * elem_size(o->id)));
}
static PyObject *sp_colptr(spmatrix *a, PyObject *u) {
    (void)u;
    return PyBytes_FromStringAndSize((const char *)a->obj->colptr, (Py_ssize_t)(((size_t)a->obj->ncols + 1) * sizeof(int_t)));
}
static PyObject *sp_rowind(spmatrix *a, PyObject *u) {
    (void)u;
    ccs *o = a->obj;
    return PyBytes_FromStringAndSize((const char *)o->rowind, (Py_ssize_t)((size_t)o->colptr[o->ncols] * sizeof(int_t)));
}
static PyObject *sp_get_size(spmatrix *a, void *c) { (void)c; return Py_BuildValue("nn", a->obj->nrows, a->obj->ncols); }

static PyMethodDef spmatrix_methods[] = {
    {"values_bytes", (PyCFunction)sp_values, METH_NOARGS, "values (float64)"},
    {"colptr_bytes", (PyCFunction)sp_colptr, METH_NOARGS, "column pointers (int64)"},
    {"rowind_bytes", (PyCFunction)sp_rowind, METH_NOARGS, "row indices (int64)"},
    {NULL, NULL, 0, NULL}};
static PyGetSetDef spmatrix_getset[] = {
    {"size", (getter)sp_get_size, NULL, "(nrows, ncols)", NULL},
    {NULL, NULL, NULL, NULL, NULL}};

/* base.spmatrix_from(values, colptr, rowind, nrows, ncols) — CCS arrays copied in */
static PyObject *py_spmatrix_from(PyObject *self, PyObject *args) {
    (void)self;
    Py_buffer v, cp, ri;
    Py_ssize_t nrows, ncols;
    if (!PyArg_ParseTuple(args, "y*y*y*nn", &v, &cp, &ri, &nrows, &ncols)) return NULL;
    spmatrix *a = NULL;
    if ((size_t)cp.len != ((size_t)ncols + 1) * sizeof(int_t) || v.len != ri.len) {
        PyErr_SetString(PyExc_ValueError, "shim: inconsistent CCS arrays");
    } else {
        int_t nnz = (int_t)(v.len / (Py_ssize_t)sizeof(double));
        a = SpMatrix_New(nrows, ncols, nnz, DOUBLE);
        if (a) {
            memcpy(a->obj->values, v.buf, (size_t)v.len);
            memcpy(a->obj->colptr, cp.buf, (size_t)cp.len);
            memcpy(a->obj->rowind, ri.buf, (size_t)ri.len);
        }
    }
    PyBuffer_Release(&v);
    PyBuffer_Release(&cp);
    PyBuffer_Release(&ri);
    return (PyObject *)a;
}

static PyTypeObject matrix_tp = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "cvxopt.base.matrix", .tp_basicsize = sizeof(matrix),
    .tp_flags = Py_TPFLAGS_DEFAULT, .tp_dealloc = (destructor)matrix_dealloc, .tp_methods = matrix_methods,
    .tp_getset = matrix_getset};
static PyTypeObject spmatrix_tp = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "cvxopt.base.spmatrix", .tp_basicsize = sizeof(spmatrix),
    .tp_flags = Py_TPFLAGS_DEFAULT, .tp_dealloc = (destructor)spmatrix_dealloc, .tp_methods = spmatrix_methods,
    .tp_getset = spmatrix_getset};

/* ---- module ---------------------------------------------------------------------------- */
static void *base_API[8];

static PyMethodDef base_functions[] = {
    {"matrix_from", py_matrix_from, METH_VARARGS, "matrix from column-major bytes"},
    {"spmatrix_from", py_spmatrix_from, METH_VARARGS, "spmatrix from CCS bytes"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef base_module = {PyModuleDef_HEAD_INIT, "base", "cvxopt.base ABI shim (test infrastructure)", -1,
                                         base_functions, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit_base(void) {
    if (PyType_Ready(&matrix_tp) < 0 || PyType_Ready(&spmatrix_tp) < 0) return NULL;
    PyObject *m = PyModule_Create(&base_module);
    if (!m) return NULL;
    Py_INCREF(&matrix_tp);
    Py_INCREF(&spmatrix_tp);
    PyModule_AddObject(m, "matrix", (PyObject *)&matrix_tp);
    PyModule_AddObject(m, "spmatrix", (PyObject *)&spmatrix_tp);
    base_API[0] = (void *)Matrix_New;
    base_API[1] = (void *)Matrix_NewFromMatrix;
    base_API[2] = (void *)Matrix_NewFromList;
    base_API[3] = (void *)Matrix_Check_func;
    base_API[4] = (void *)SpMatrix_New;
    base_API[5] = (void *)SpMatrix_NewFromSpMatrix;
    base_API[6] = (void *)SpMatrix_NewFromIJV;
    base_API[7] = (void *)SpMatrix_Check_func;
    PyObject *cap = PyCapsule_New((void *)base_API, "base_API", NULL);
    if (cap) PyModule_AddObject(m, "_C_API", cap);
    return m;
}
