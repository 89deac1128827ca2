// bvh8_build.cpp -- host builder of the 8-wide compressed BVH (see bvh8.h).
//
// Replaces BVHAccel's constructor (accelerators/bvh.cpp:183-225: SAH
// recursiveBuild :236-402, flattenBVHTree :640-658).  The tree topology is
// not part of the parity contract -- closest hits are decided by the exact
// watertight triangle test and are topology independent (DESIGN.md,
// "Traversal order and ties") -- so the build is designed for the GPU
// traversal kernel, not to mimic the reference's binary tree:
//   1. binary BVH by binned SAH (16 bins), leaves of <= 3 triangles, built
//      top-down with one std::thread per large subtree;
//   2. greedy collapse to 8 children per node (open the child with the
//      largest surface area until 8 slots are used);
//   3. children placed in octant-ordered slots (greedy assignment on
//      dot(centroid offset, octant direction)) so that visiting slots by
//      (slot XOR ray octant) approximates front-to-back order;
//   4. child boxes quantised to 8 bits per coordinate on a power-of-two grid
//      anchored one cell below the node's min corner, rounded outward by one
//      extra cell on every side (the traversal kernel's fused slab arithmetic may be
//      off by up to half a cell; a second cell of slack costs 7 % more node visits).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>

#include "bvh8.h"

namespace b200pt {
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int a = 0; a < 3; ++a) {
            lo[a] = INFINITY;
            hi[a] = -INFINITY;
        }
    }
    void grow(const Box &b) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    void grow(const float *p) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], p[a]);
            hi[a] = std::max(hi[a], p[a]);
        }
    }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0) || !(dy >= 0) || !(dz >= 0)) return 0.f;
        return dx * dy + dy * dz + dz * dx;
    }
};

struct Node2 {
    Box box;
    int32_t left, right;   // children (inner) or -1
    int32_t first, count;  // leaf range in idx[] (count > 0 <=> binary leaf = one triangle)
    int32_t ntri;          // triangles in the subtree: idx[first .. first + ntri)
};

struct Builder2 {
    const Box *tbox;
    const float *cent;  // 3 per triangle
    std::vector<int32_t> idx;
    std::vector<Node2> nodes;
    std::atomic<int32_t> next_node{0};
    std::atomic<int> threads_free{0};

    int32_t alloc() { return next_node.fetch_add(1); }

    void build(int32_t node, int32_t first, int32_t count) {
        Node2 &n = nodes[node];
        Box b, cb;
        b.reset();
        cb.reset();
        for (int32_t i = first; i < first + count; ++i) {
            b.grow(tbox[idx[i]]);
            cb.grow(cent + 3 * (size_t)idx[i]);
        }
        n.box = b;
        n.left = n.right = -1;
        n.first = first;
        n.count = 0;
        n.ntri = count;
        if (count == 1) {
            n.count = 1;
            return;
        }
        // binned SAH over the centroid bounds
        const int NB = 16;
        int bestAxis = -1, bestBin = -1;
        float bestCost = INFINITY;
        for (int axis = 0; axis < 3; ++axis) {
            float c0 = cb.lo[axis], c1 = cb.hi[axis];
            if (!(c1 > c0)) continue;
            float scale = NB / (c1 - c0);
            Box bb[NB];
            int bc[NB];
            for (int i = 0; i < NB; ++i) {
                bb[i].reset();
                bc[i] = 0;
            }
            for (int32_t i = first; i < first + count; ++i) {
                int t = idx[i];
                int bi = std::min(NB - 1, std::max(0, (int)((cent[3 * (size_t)t + axis] - c0) * scale)));
                bb[bi].grow(tbox[t]);
                bc[bi]++;
            }
            float rightArea[NB];
            int rightCount[NB];
            Box acc;
            acc.reset();
            int cnt = 0;
            for (int i = NB - 1; i > 0; --i) {
                acc.grow(bb[i]);
                cnt += bc[i];
                rightArea[i] = acc.half_area();
                rightCount[i] = cnt;
            }
            acc.reset();
            cnt = 0;
            for (int i = 0; i < NB - 1; ++i) {
                acc.grow(bb[i]);
                cnt += bc[i];
                if (cnt == 0 || rightCount[i + 1] == 0) continue;
                float cost = acc.half_area() * cnt + rightArea[i + 1] * rightCount[i + 1];
                if (cost < bestCost) {
                    bestCost = cost;
                    bestAxis = axis;
                    bestBin = i;
                }
            }
        }
        int32_t mid;
        if (bestAxis >= 0) {
            float c0 = cb.lo[bestAxis], scale = NB / (cb.hi[bestAxis] - c0);
            const float *cc = cent;
            int ax = bestAxis, bbn = bestBin;
            auto it = std::partition(idx.begin() + first, idx.begin() + first + count, [=](int32_t t) {
                int bi = std::min(NB - 1, std::max(0, (int)((cc[3 * (size_t)t + ax] - c0) * scale)));
                return bi <= bbn;
            });
            mid = (int32_t)(it - idx.begin());
        } else {
            mid = first + count / 2;  // all centroids coincide: split the range in half
        }
        if (mid == first || mid == first + count) mid = first + count / 2;
        int32_t l = alloc(), r = alloc();
        nodes[node].left = l;
        nodes[node].right = r;
        int32_t lc = mid - first, rc = count - lc;
        if (count > 65536 && threads_free.fetch_sub(1) > 0) {
            std::thread th([this, l, first, lc]() { build(l, first, lc); });
            build(r, mid, rc);
            th.join();
            threads_free.fetch_add(1);
        } else {
            if (count > 65536) threads_free.fetch_add(1);
            build(l, first, lc);
            build(r, mid, rc);
        }
    }
};

struct Child {
    int32_t node2;  // binary node
    Box box;
};

}  // namespace

void build_bvh8(const float *vertices, int64_t n_tris, const int32_t *material_id, const int32_t *light_id,
                const uint8_t *flip, const uint8_t *degenerate, int n_threads, Bvh8 *out) {
    out->nodes.clear();
    out->tris.clear();
    out->prim_to_tri.assign((size_t)n_tris, 0xffffffffu);
    out->n_in_leaves = 0;
    out->max_depth = 0;

    std::vector<Box> tbox((size_t)n_tris);
    std::vector<float> cent(3 * (size_t)n_tris);
    Builder2 b2;
    b2.idx.reserve((size_t)n_tris);
    for (int64_t i = 0; i < n_tris; ++i) {
        Box b;
        b.reset();
        const float *v = vertices + 9 * i;
        b.grow(v);
        b.grow(v + 3);
        b.grow(v + 6);
        tbox[i] = b;
        bool finite = true;
        for (int a = 0; a < 3; ++a) {
            cent[3 * i + a] = 0.5f * b.lo[a] + 0.5f * b.hi[a];
            finite = finite && std::isfinite(b.lo[a]) && std::isfinite(b.hi[a]);
        }
        if (!(degenerate && degenerate[i]) && finite) b2.idx.push_back((int32_t)i);
    }
    auto make_tri = [&](int64_t i) {
        TriRecord t;
        const float *v = vertices + 9 * i;
        memcpy(t.p0, v, 12);
        memcpy(t.p1, v + 3, 12);
        memcpy(t.p2, v + 6, 12);
        t.prim = (uint32_t)i;
        t.mat_flags = (uint32_t)(material_id ? material_id[i] : 0) | ((flip && flip[i]) ? 0x10000u : 0u) |
                      ((degenerate && degenerate[i]) ? 0x20000u : 0u);
        t.light = light_id ? light_id[i] : -1;
        return t;
    };
    const int32_t nLeafTris = (int32_t)b2.idx.size();
    out->tris.reserve((size_t)n_tris);

    if (nLeafTris > 0) {
        b2.tbox = tbox.data();
        b2.cent = cent.data();
        b2.nodes.resize(2 * (size_t)nLeafTris);
        b2.threads_free = std::max(0, n_threads - 1);
        int32_t root = b2.alloc();
        const bool trace = getenv("B200PT_BUILD_TRACE") != nullptr;
        auto tnow = []() { return std::chrono::steady_clock::now(); };
        auto tp0 = tnow();
        b2.build(root, 0, nLeafTris);
        auto tp1 = tnow();

        // ---- SAH-optimal collapse (Ylitie et al. 2017, section 3.1): cost[n][i-1] is the cheapest
        // way to represent the binary subtree n with at most i roots (i = 1..7); a single root is
        // either a leaf child (<= 3 triangles) or an 8-wide node whose children come from
        // distributing the two binary children over 8 slots.
        const int32_t nNodes2 = b2.next_node.load();
        float cNode = 1.0f, cPrim = 1.0f;  // a watertight triangle test costs about as many instructions as a node
        if (const char *e = getenv("B200PT_SAH_CNODE")) cNode = (float)atof(e);
        if (const char *e = getenv("B200PT_SAH_CPRIM")) cPrim = (float)atof(e);
        std::vector<float> cost((size_t)nNodes2 * 7);
        std::vector<uint8_t> dec((size_t)nNodes2 * 7);   // i=1: 0 leaf / 1 inner ; i>=2: 0 = reuse i-1, k = left gets k roots
        std::vector<uint8_t> split8((size_t)nNodes2);    // left share when the node becomes an 8-wide node
        for (int32_t n = nNodes2 - 1; n >= 0; --n) {
            const Node2 &nd = b2.nodes[n];
            float *c = &cost[(size_t)n * 7];
            uint8_t *d = &dec[(size_t)n * 7];
            const float A = nd.box.half_area();
            const float leafCost = nd.ntri <= 3 ? A * nd.ntri * cPrim : INFINITY;
            if (nd.count > 0) {  // single triangle
                for (int i = 0; i < 7; ++i) {
                    c[i] = leafCost;
                    d[i] = 0;
                }
                split8[n] = 0;
                continue;
            }
            const float *cl = &cost[(size_t)nd.left * 7], *cr = &cost[(size_t)nd.right * 7];
            auto distribute = [&](int j, int *bestK) {
                float best = INFINITY;
                *bestK = 1;
                for (int k = 1; k < j; ++k) {
                    if (k > 7 || j - k > 7) continue;
                    float v = cl[k - 1] + cr[j - k - 1];
                    if (v < best) {
                        best = v;
                        *bestK = k;
                    }
                }
                return best;
            };
            int k8;
            const float innerCost = distribute(8, &k8) + A * cNode;
            split8[n] = (uint8_t)k8;
            if (leafCost <= innerCost) {
                c[0] = leafCost;
                d[0] = 0;
            } else {
                c[0] = innerCost;
                d[0] = 1;
            }
            for (int i = 2; i <= 7; ++i) {
                int k;
                float v = distribute(i, &k);
                if (v < c[i - 2]) {
                    c[i - 1] = v;
                    d[i - 1] = (uint8_t)k;
                } else {
                    c[i - 1] = c[i - 2];
                    d[i - 1] = 0;
                }
            }
        }
        auto tp2 = tnow();
        // children of a wide node: expand binary node `n` into at most `budget` roots
        struct Collector {
            const Builder2 &b2;
            const std::vector<uint8_t> &dec;
            Child *ch;
            int k;
            void collect(int32_t n, int budget) {
                const Node2 &nd = b2.nodes[n];
                int i = budget;
                while (i >= 2 && dec[(size_t)n * 7 + i - 1] == 0) --i;  // reuse the (i-1)-root solution
                if (i == 1 || nd.count > 0) {
                    ch[k++] = {n, nd.box};
                    return;
                }
                const int kl = dec[(size_t)n * 7 + i - 1];
                collect(nd.left, kl);
                collect(nd.right, i - kl);
            }
        };

        // ---- emit 8-wide nodes breadth-first so siblings are contiguous
        struct Pending {
            int32_t node2;
            uint32_t wide;
            int depth;
        };
        std::vector<Pending> queue;
        out->nodes.reserve((size_t)nLeafTris / 3 + 16);
        out->nodes.push_back(Bvh8Node());
        queue.push_back({root, 0u, 1});
        for (size_t qi = 0; qi < queue.size(); ++qi) {
            Pending cur = queue[qi];
            out->max_depth = std::max(out->max_depth, cur.depth);
            Child ch[8];
            int k = 0;
            const Node2 &rn = b2.nodes[cur.node2];
            if (rn.count > 0 || (cur.node2 == root && dec[(size_t)root * 7] == 0)) {
                ch[k++] = {cur.node2, rn.box};  // the whole scene is one leaf child of the root
            } else {
                Collector col{b2, dec, ch, 0};
                col.collect(rn.left, split8[cur.node2]);
                col.collect(rn.right, 8 - split8[cur.node2]);
                k = col.k;
            }
            // node bounds and centroid
            Box nb;
            nb.reset();
            for (int i = 0; i < k; ++i) nb.grow(ch[i].box);
            float nc[3];
            for (int a = 0; a < 3; ++a) nc[a] = 0.5f * nb.lo[a] + 0.5f * nb.hi[a];
            // octant-ordered slot assignment (greedy)
            int slotOf[8];
            bool slotUsed[8] = {false, false, false, false, false, false, false, false};
            bool childDone[8] = {false, false, false, false, false, false, false, false};
            float cost[8][8];
            for (int i = 0; i < k; ++i) {
                float cc[3];
                for (int a = 0; a < 3; ++a) cc[a] = 0.5f * ch[i].box.lo[a] + 0.5f * ch[i].box.hi[a] - nc[a];
                for (int s = 0; s < 8; ++s)
                    cost[i][s] = ((s & 1) ? cc[0] : -cc[0]) + ((s & 2) ? cc[1] : -cc[1]) + ((s & 4) ? cc[2] : -cc[2]);
            }
            for (int it = 0; it < k; ++it) {
                int bi = -1, bs = -1;
                float bc = -INFINITY;
                for (int i = 0; i < k; ++i) {
                    if (childDone[i]) continue;
                    for (int s = 0; s < 8; ++s)
                        if (!slotUsed[s] && (cost[i][s] > bc || bi < 0)) {
                            bc = cost[i][s];
                            bi = i;
                            bs = s;
                        }
                }
                childDone[bi] = true;
                slotUsed[bs] = true;
                slotOf[bi] = bs;
            }
            int childAt[8];
            for (int s = 0; s < 8; ++s) childAt[s] = -1;
            for (int i = 0; i < k; ++i) childAt[slotOf[i]] = i;

            Bvh8Node node;
            memset(&node, 0, sizeof(node));
            // quantisation grid
            float scale[3];
            for (int a = 0; a < 3; ++a) {
                float ext = nb.hi[a] - nb.lo[a];
                float mag = std::max(std::fabs(nb.lo[a]), std::fabs(nb.hi[a]));
                float need = std::max(ext / 253.f, std::max(mag * 0x1p-18f, 1e-30f));
                int e;
                float m = std::frexp(need, &e);  // need = m * 2^e, m in [0.5,1)
                if (m == 0.5f) e -= 1;           // exact power of two
                int be = std::min(254, std::max(1, e + 127));
                node.e[a] = (uint8_t)be;
                uint32_t bits = (uint32_t)be << 23;
                memcpy(&scale[a], &bits, 4);
                node.p[a] = nb.lo[a] - scale[a];
            }
            node.child_base = (uint32_t)out->nodes.size();
            node.tri_base = (uint32_t)out->tris.size();
            int triOffset = 0;
            for (int s = 0; s < 8; ++s) {
                int i = childAt[s];
                if (i < 0) continue;
                const Node2 &c = b2.nodes[ch[i].node2];
                const bool isLeaf = c.count > 0 || dec[(size_t)ch[i].node2 * 7] == 0;
                for (int a = 0; a < 3; ++a) {
                    float lo = std::floor((ch[i].box.lo[a] - node.p[a]) / scale[a]) - 1.f;
                    float hi = std::ceil((ch[i].box.hi[a] - node.p[a]) / scale[a]) + 1.f;
                    int qlo = (int)std::min(255.f, std::max(0.f, lo));
                    int qhi = (int)std::min(255.f, std::max(0.f, hi));
                    // make sure the decode the kernel performs (p + q*scale in float) is conservative
                    while (qlo > 0 && node.p[a] + (float)qlo * scale[a] > ch[i].box.lo[a]) --qlo;
                    while (qhi < 255 && node.p[a] + (float)qhi * scale[a] < ch[i].box.hi[a]) ++qhi;
                    node.qlo[a][s] = (uint8_t)qlo;
                    node.qhi[a][s] = (uint8_t)qhi;
                }
                if (isLeaf) {
                    const uint8_t unary[4] = {0, 1, 3, 7};
                    node.meta[s] = (uint8_t)((unary[c.ntri] << 5) | triOffset);
                    for (int t = 0; t < c.ntri; ++t) {
                        int32_t tri = b2.idx[c.first + t];
                        out->prim_to_tri[tri] = (uint32_t)out->tris.size();
                        out->tris.push_back(make_tri(tri));
                    }
                    triOffset += c.ntri;
                } else {
                    node.imask |= (uint8_t)(1u << s);
                    node.meta[s] = (uint8_t)((1u << 5) | (24 + s));
                    queue.push_back({ch[i].node2, (uint32_t)out->nodes.size(), cur.depth + 1});
                    out->nodes.push_back(Bvh8Node());
                }
            }
            out->nodes[cur.wide] = node;
        }
        out->n_in_leaves = (uint32_t)out->tris.size();
        if (trace)
            fprintf(stderr, "bvh8 build: binary SAH %.3f s, collapse DP %.3f s, emit %.3f s\n",
                    std::chrono::duration<double>(tp1 - tp0).count(), std::chrono::duration<double>(tp2 - tp1).count(),
                    std::chrono::duration<double>(tnow() - tp2).count());
    } else {
        // empty scene: a root with no children
        Bvh8Node node;
        memset(&node, 0, sizeof(node));
        node.e[0] = node.e[1] = node.e[2] = 127;
        out->nodes.push_back(node);
        out->max_depth = 1;
    }
    // triangles that can never be hit still need records (an area light may sit on one)
    for (int64_t i = 0; i < n_tris; ++i)
        if (out->prim_to_tri[i] == 0xffffffffu) {
            out->prim_to_tri[i] = (uint32_t)out->tris.size();
            out->tris.push_back(make_tri(i));
        }
}

int64_t validate_bvh8(const Bvh8 &bvh) {
    int64_t bad = 0;
    std::vector<Box> nodeBox(bvh.nodes.size());
    // decoded slot boxes must contain what they refer to
    auto slot_box = [](const Bvh8Node &n, int s) {
        Box b;
        for (int a = 0; a < 3; ++a) {
            uint32_t bits = (uint32_t)n.e[a] << 23;
            float sc;
            memcpy(&sc, &bits, 4);
            b.lo[a] = n.p[a] + (float)n.qlo[a][s] * sc;
            b.hi[a] = n.p[a] + (float)n.qhi[a][s] * sc;
        }
        return b;
    };
    // exact box of each wide node = union of decoded slots is an over-estimate; compute true
    // content bounds bottom-up (children have larger indices than their parents)
    std::vector<Box> content(bvh.nodes.size());
    for (int64_t ni = (int64_t)bvh.nodes.size() - 1; ni >= 0; --ni) {
        const Bvh8Node &n = bvh.nodes[ni];
        Box cb;
        cb.reset();
        uint32_t inner = 0;
        for (int s = 0; s < 8; ++s) {
            uint8_t m = n.meta[s];
            if (m == 0) continue;
            Box sb = slot_box(n, s);
            Box got;
            got.reset();
            if (n.imask & (1u << s)) {
                uint32_t child = n.child_base + inner++;
                if (child >= bvh.nodes.size() || child <= (uint32_t)ni) {
                    ++bad;
                    continue;
                }
                got = content[child];
            } else {
                int cnt = __builtin_popcount(m >> 5), off = m & 31;
                for (int t = 0; t < cnt; ++t) {
                    const TriRecord &tr = bvh.tris[n.tri_base + off + t];
                    got.grow(tr.p0);
                    got.grow(tr.p1);
                    got.grow(tr.p2);
                }
            }
            for (int a = 0; a < 3; ++a)
                if (!(sb.lo[a] <= got.lo[a]) || !(sb.hi[a] >= got.hi[a])) ++bad;
            cb.grow(got);
        }
        content[ni] = cb;
    }
    return bad;
}

}  // namespace b200pt
