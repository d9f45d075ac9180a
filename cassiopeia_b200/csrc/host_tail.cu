// Host-side passes of the array-based solve tail (cassiopeia_b200/tree_tail.py): the
// three per-node loops that do not vectorise (depth-first preorder, heights, successor lists after edge
// contraction).  Plain host code -- no device work -- kept in this library so that the Python tail makes
// three C calls instead of walking 10^5 nodes in the interpreter.  What they compute, and the reference
// passes they stand for, is described in tree_tail.py and include/cas_b200.h.
#include <cstdint>
#include <vector>

#include "../../include/cas_b200.h"
#include "common.cuh"

extern "C" {

int64_t cas_tail_preorder(const int64_t *child0, const int64_t *child1, int64_t n_nodes, int64_t root,
                          int64_t *order, int64_t *depth) {
  if (!child0 || !child1 || !order || !depth || n_nodes <= 0 || root < 0 || root >= n_nodes) {
    cas::set_error("cas_tail_preorder: invalid arguments");
    return CAS_ERR_INVALID;
  }
  std::vector<int64_t> stack;
  stack.reserve(256);
  stack.push_back(root);
  depth[root] = 0;
  int64_t count = 0;
  while (!stack.empty()) {
    const int64_t v = stack.back();
    stack.pop_back();
    if (count >= n_nodes) {
      cas::set_error("cas_tail_preorder: the child arrays do not describe a tree (a node is reached twice)");
      return CAS_ERR_INVALID;
    }
    order[count++] = v;
    const int64_t x = child0[v], y = child1[v];
    if (x >= n_nodes || y >= n_nodes) {
      cas::set_error("cas_tail_preorder: child index out of range");
      return CAS_ERR_INVALID;
    }
    if (x >= 0) {
      const int64_t d = depth[v] + 1;
      if (y >= 0) {
        depth[y] = d;
        stack.push_back(y);
      }
      depth[x] = d;
      stack.push_back(x);  // popped first: first child first
    }
  }
  return count;
}

int cas_tail_heights(const int64_t *order, int64_t n_order, const int64_t *child0, const int64_t *child1,
                     int64_t *height) {
  if (!order || !child0 || !child1 || !height || n_order < 0) {
    cas::set_error("cas_tail_heights: invalid arguments");
    return CAS_ERR_INVALID;
  }
  for (int64_t p = n_order - 1; p >= 0; --p) {  // children precede parents in reverse preorder
    const int64_t v = order[p];
    const int64_t x = child0[v], y = child1[v];
    int64_t h = 0;
    if (x >= 0) {
      h = height[x];
      if (y >= 0 && height[y] > h) h = height[y];
      h += 1;
    }
    height[v] = h;
  }
  return CAS_OK;
}

int cas_tail_successor_lists(const int64_t *order, int64_t n_order, const int64_t *child0, const int64_t *child1,
                             const uint8_t *removed, int64_t *head, int64_t *tail, int64_t *next) {
  if (!order || !child0 || !child1 || !removed || !head || !tail || !next || n_order < 0) {
    cas::set_error("cas_tail_successor_lists: invalid arguments");
    return CAS_ERR_INVALID;
  }
  for (int64_t p = n_order - 1; p >= 0; --p) {  // children before parents
    const int64_t v = order[p];
    const int64_t kids[2] = {child0[v], child1[v]};
    int64_t h = -1, t = -1;
    if (kids[0] >= 0) {
      for (int c = 0; c < 2; ++c) {  // kept children, in order
        const int64_t k = kids[c];
        if (k < 0 || removed[k]) continue;
        if (h < 0) h = t = k;
        else { next[t] = k; t = k; }
      }
      for (int c = 0; c < 2; ++c) {  // then the lists of the spliced children
        const int64_t k = kids[c];
        if (k < 0 || !removed[k] || head[k] < 0) continue;
        if (h < 0) { h = head[k]; t = tail[k]; }
        else { next[t] = head[k]; t = tail[k]; }
      }
    }
    head[v] = h;
    tail[v] = t;
  }
  return CAS_OK;
}

}  // extern "C"
