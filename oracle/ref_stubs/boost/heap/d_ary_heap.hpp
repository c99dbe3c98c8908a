// ORACLE / TEST INFRASTRUCTURE — stand-in for boost::heap::d_ary_heap (Boost is not in the image), just enough for the
// reference's graph_search.h to compile: a mutable binary heap with handles. Only needed so that
// GraphSearch::getDensePath (graph_search.cpp:119-176) can be called from the compiled reference; the A* / JPS
// search that uses the heap is out of scope and not pinned (tie order inside a heap is implementation-defined).
#pragma once
#include <cstddef>
#include <memory>
#include <utility>
#include <vector>

namespace boost { namespace heap {
template <bool B> struct mutable_ {};
template <int N> struct arity {};
template <class C> struct compare { typedef C type; };

template <class T, class M, class A, class CmpTag>
class d_ary_heap {
    struct Node { T value; std::size_t pos; };
    typedef typename CmpTag::type Cmp;
    std::vector<std::shared_ptr<Node>> h_;
    Cmp cmp_;   // cmp_(a, b): a has LOWER priority than b (boost max-heap convention)
    void place(std::size_t i, std::shared_ptr<Node> n) { n->pos = i; h_[i] = std::move(n); }
    void up(std::size_t i) {
        std::shared_ptr<Node> n = h_[i];
        while (i > 0) {
            const std::size_t p = (i - 1) / 2;
            if (!cmp_(h_[p]->value, n->value)) break;
            place(i, h_[p]);
            i = p;
        }
        place(i, n);
    }
    void down(std::size_t i) {
        std::shared_ptr<Node> n = h_[i];
        const std::size_t sz = h_.size();
        for (;;) {
            std::size_t c = 2 * i + 1;
            if (c >= sz) break;
            if (c + 1 < sz && cmp_(h_[c]->value, h_[c + 1]->value)) c++;
            if (!cmp_(n->value, h_[c]->value)) break;
            place(i, h_[c]);
            i = c;
        }
        place(i, n);
    }
public:
    struct handle_type { std::shared_ptr<Node> n; };
    bool empty() const { return h_.empty(); }
    std::size_t size() const { return h_.size(); }
    void clear() { h_.clear(); }
    const T& top() const { return h_.front()->value; }
    handle_type push(const T& v) {
        auto n = std::make_shared<Node>(Node{v, h_.size()});
        h_.push_back(n);
        up(h_.size() - 1);
        return handle_type{n};
    }
    void pop() {
        if (h_.size() > 1) { place(0, h_.back()); h_.pop_back(); down(0); }
        else h_.clear();
    }
    void increase(const handle_type& h) { up(h.n->pos); }
    void decrease(const handle_type& h) { down(h.n->pos); }
    void update(const handle_type& h) { up(h.n->pos); down(h.n->pos); }
};
}}  // namespace boost::heap
