/* CPython extension `cassiopeia_b200._graphfill`: fills the three dictionaries of an empty networkx DiGraph
 * (node -> attrs, u -> {v -> attrs}, v -> {u -> attrs}, the edge attribute dict shared between the last two)
 * from the arrays of cassiopeia_b200/tree_tail.py.  It is the last per-node loop of the solve tail
 * (SURVEY §8(f) rank 1): the result is exactly what tree_tail.build_final_tree's interpreter loop builds --
 * node attributes as the reference's populate_tree leaves them (`character_states` = [], `time`;
 * cassiopeia/data/CassiopeiaTree.py:174-203) and `length` on every edge -- but 10^5 nodes take 40 ms
 * instead of 250 ms.  Host code only; tree_tail.py falls back to its own loop when the module is absent. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>

static int get_i64(PyObject *obj, Py_buffer *view, const char *what) {
  if (PyObject_GetBuffer(obj, view, PyBUF_CONTIG_RO | PyBUF_FORMAT) != 0) return -1;
  if (view->itemsize != 8 || view->ndim != 1 || view->format == NULL ||
      (view->format[0] != 'l' && view->format[0] != 'q')) {
    PyErr_Format(PyExc_TypeError, "%s must be a contiguous 1-d int64 array", what);
    PyBuffer_Release(view);
    return -1;
  }
  return 0;
}

/* fill(node, succ, pred, labels, kept_order, head, next, depth) -> None */
static PyObject *graphfill_fill(PyObject *self, PyObject *args) {
  PyObject *node, *succ, *pred, *labels, *o_order, *o_head, *o_next, *o_depth;
  if (!PyArg_ParseTuple(args, "O!O!O!O!OOOO", &PyDict_Type, &node, &PyDict_Type, &succ, &PyDict_Type, &pred,
                        &PyList_Type, &labels, &o_order, &o_head, &o_next, &o_depth))
    return NULL;
  Py_buffer b_order, b_head, b_next, b_depth;
  if (get_i64(o_order, &b_order, "kept_order") != 0) return NULL;
  if (get_i64(o_head, &b_head, "head") != 0) { PyBuffer_Release(&b_order); return NULL; }
  if (get_i64(o_next, &b_next, "next") != 0) { PyBuffer_Release(&b_order); PyBuffer_Release(&b_head); return NULL; }
  if (get_i64(o_depth, &b_depth, "depth") != 0) {
    PyBuffer_Release(&b_order); PyBuffer_Release(&b_head); PyBuffer_Release(&b_next);
    return NULL;
  }
  const int64_t *order = (const int64_t *)b_order.buf, *head = (const int64_t *)b_head.buf;
  const int64_t *next = (const int64_t *)b_next.buf, *depth = (const int64_t *)b_depth.buf;
  const Py_ssize_t n_kept = b_order.len / 8, M = b_head.len / 8;
  PyObject **succ_of = NULL, **pred_of = NULL;
  PyObject *k_states = NULL, *k_time = NULL, *k_length = NULL;
  int ok = 0;
  Py_ssize_t base = 0;

  if (b_next.len / 8 != M || b_depth.len / 8 != M || PyList_GET_SIZE(labels) < M) {
    PyErr_SetString(PyExc_ValueError, "head, next, depth and labels must cover the same nodes");
    goto done;
  }
  succ_of = (PyObject **)calloc((size_t)(M > 0 ? M : 1), sizeof(PyObject *));
  pred_of = (PyObject **)calloc((size_t)(M > 0 ? M : 1), sizeof(PyObject *));
  if (!succ_of || !pred_of) { PyErr_NoMemory(); goto done; }
  k_states = PyUnicode_InternFromString("character_states");
  k_time = PyUnicode_InternFromString("time");
  k_length = PyUnicode_InternFromString("length");
  if (!k_states || !k_time || !k_length) goto done;

  /* nodes, in the given order (= insertion order of the reference's graph) */
  base = PyDict_GET_SIZE(node);
  for (Py_ssize_t p = 0; p < n_kept; ++p) {
    const int64_t v = order[p];
    if (v < 0 || v >= M) { PyErr_SetString(PyExc_ValueError, "node id out of range"); goto done; }
    PyObject *lv = PyList_GET_ITEM(labels, v); /* borrowed */
    PyObject *attrs = PyDict_New(), *states = PyList_New(0), *t = PyLong_FromLongLong(depth[v]);
    PyObject *s = PyDict_New(), *q = PyDict_New();
    int bad = !attrs || !states || !t || !s || !q;
    if (!bad) bad = PyDict_SetItem(attrs, k_states, states) != 0 || PyDict_SetItem(attrs, k_time, t) != 0;
    if (!bad) bad = PyDict_SetItem(node, lv, attrs) != 0 || PyDict_SetItem(succ, lv, s) != 0 ||
                    PyDict_SetItem(pred, lv, q) != 0;
    Py_XDECREF(attrs); Py_XDECREF(states); Py_XDECREF(t);
    /* the dictionaries now hold the references that keep s and q alive */
    Py_XDECREF(s); Py_XDECREF(q);
    if (bad) goto done;
    if (PyDict_GET_SIZE(node) != base + p + 1) {
      /* a repeated name replaced an entry: the borrowed pointers below would dangle */
      PyErr_SetString(PyExc_ValueError, "node names must be unique");
      goto done;
    }
    succ_of[v] = s;
    pred_of[v] = q;
  }
  /* edges: every kept node's successor list, in list order */
  for (Py_ssize_t p = 0; p < n_kept; ++p) {
    const int64_t v = order[p];
    PyObject *lv = PyList_GET_ITEM(labels, v);
    int64_t c = head[v];
    Py_ssize_t guard = 0;
    while (c >= 0) {
      if (c >= M || !pred_of[c] || ++guard > M) {
        PyErr_SetString(PyExc_ValueError, "successor list refers to a node that was not kept");
        goto done;
      }
      PyObject *attrs = PyDict_New(), *len = PyLong_FromLongLong(depth[c] - depth[v]);
      int bad = !attrs || !len || PyDict_SetItem(attrs, k_length, len) != 0 ||
                PyDict_SetItem(succ_of[v], PyList_GET_ITEM(labels, c), attrs) != 0 ||
                PyDict_SetItem(pred_of[c], lv, attrs) != 0;
      Py_XDECREF(attrs); Py_XDECREF(len);
      if (bad) goto done;
      c = next[c];
    }
  }
  ok = 1;
done:
  free(succ_of); free(pred_of);
  Py_XDECREF(k_states); Py_XDECREF(k_time); Py_XDECREF(k_length);
  PyBuffer_Release(&b_order); PyBuffer_Release(&b_head); PyBuffer_Release(&b_next); PyBuffer_Release(&b_depth);
  if (!ok) return NULL;
  Py_RETURN_NONE;
}

static PyMethodDef graphfill_methods[] = {
    {"fill", graphfill_fill, METH_VARARGS,
     "fill(node, succ, pred, labels, kept_order, head, next, depth): populate the dictionaries of an empty "
     "networkx DiGraph from int64 arrays"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef graphfill_module = {PyModuleDef_HEAD_INIT, "_graphfill",
                                              "dictionary filler of the array-based solve tail", -1, graphfill_methods};

PyMODINIT_FUNC PyInit__graphfill(void) { return PyModule_Create(&graphfill_module); }
