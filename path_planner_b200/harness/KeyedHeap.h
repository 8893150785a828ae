// KeyedHeap.h -- the heap operations SamplingBasedPlanner::expand performs on m_Samples
// (SamplingBasedPlanner.cpp:85 std::make_heap, :93 std::pop_heap, comparator "farther from the source vertex is
// lower priority", :36-40), on a (key, index) representation: key[i] = distance of logical element i to the source,
// idx[i] = which sample it is.  The reference's comparator recomputes two square roots per comparison and the heap
// moves 40-byte States; here every distance is computed once per expansion and 12 bytes move.
//
// The routines are libstdc++'s algorithms (bits/stl_heap.h: __make_heap, __adjust_heap, __push_heap, __pop_heap)
// restated step for step, so the ARRANGEMENT they leave behind is the one std::make_heap / std::pop_heap leave in the
// reference built with the same standard library -- which matters only when two samples are at exactly the same
// distance (pop order among ties), and is checked against the real std:: functions, ties included, by
// tests/test_harness_host_logic.py::test_keyed_heap_matches_libstdcxx.
#ifndef PPE_KEYED_HEAP_H
#define PPE_KEYED_HEAP_H

#include <cstddef>
#include <cstdint>

namespace ppe_heap {

// comp(a, b) of the reference: a is lower priority than b  <=>  key[a] > key[b]
inline void push_heap_(double* key, uint32_t* idx, std::ptrdiff_t hole, std::ptrdiff_t top, double vkey, uint32_t vidx) {
    std::ptrdiff_t parent = (hole - 1) / 2;
    while (hole > top && key[parent] > vkey) {
        key[hole] = key[parent];
        idx[hole] = idx[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    key[hole] = vkey;
    idx[hole] = vidx;
}

inline void adjust_heap_(double* key, uint32_t* idx, std::ptrdiff_t hole, std::ptrdiff_t len, double vkey, uint32_t vidx) {
    const std::ptrdiff_t top = hole;
    std::ptrdiff_t second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (key[second] > key[second - 1]) second--;
        key[hole] = key[second];
        idx[hole] = idx[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        key[hole] = key[second - 1];
        idx[hole] = idx[second - 1];
        hole = second - 1;
    }
    push_heap_(key, idx, hole, top, vkey, vidx);
}

// std::make_heap(first, first + len, comp)
inline void make_heap(double* key, uint32_t* idx, std::ptrdiff_t len) {
    if (len < 2) return;
    std::ptrdiff_t parent = (len - 2) / 2;
    for (;;) {
        const double vkey = key[parent];
        const uint32_t vidx = idx[parent];
        adjust_heap_(key, idx, parent, len, vkey, vidx);
        if (parent == 0) return;
        parent--;
    }
}

// std::pop_heap(first, first + len, comp): the top moves to position len - 1, [0, len - 1) is a heap again
inline void pop_heap(double* key, uint32_t* idx, std::ptrdiff_t len) {
    if (len < 2) return;
    const std::ptrdiff_t last = len - 1;
    const double vkey = key[last];
    const uint32_t vidx = idx[last];
    key[last] = key[0];
    idx[last] = idx[0];
    adjust_heap_(key, idx, 0, last, vkey, vidx);
}

} // namespace ppe_heap

#endif
