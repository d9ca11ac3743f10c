// bvh_wide.cu - warp-cooperative closest point on the static target surface.
//
// Replaces target.operations.closestPointOnSurface (NonRigidIcpProposal.scala:97, IndependentPointDistanceEvaluator.scala:43,
// CollectiveAverage...Evaluator.scala:45, RegistrationComparison.scala:35) for the STATIC target; the per-thread kernel in
// bvh.cu keeps the refitted (per-chain) structures.
//
// Structure: the device-built binary LBVH (bvh.cu) collapsed into an L-ary tree (L = 4 or 8). A wide node is L child
// records of 32 bytes {box lo xyz, box hi xyz (FP32, conservative), reference}; a child is another wide node or a leaf
// cluster of up to L triangles that are CONTIGUOUS in Morton order (a subtree of the LBVH covers a contiguous range of
// sorted leaves). Nodes are numbered breadth first, so a prefix of the array is the top of the tree.
//
// Traversal: L lanes per query (32 / L queries per warp). At a wide node every lane tests ONE child box, the sub-warp
// picks the nearest hit with shuffles and pushes the others on ONE stack per query in shared memory; at a leaf cluster
// every lane runs the exact FP64 point-triangle routine on ONE triangle. A query therefore takes log_L instead of log_2
// dependent steps, and the lanes of a query never diverge from each other. The node array (a few tens of KB for a
// femur-sized target) is staged in shared memory by persistent CTAs; deeper nodes of larger trees come from global
// memory. Exactness: subtrees are only pruned when their conservative lower bound exceeds the best distance, ties go
// to the lowest triangle index, so the result equals a brute-force FP64 search - the same contract as bvh.cu.
#include <algorithm>
#include <queue>
#include <vector>

#include "bvh_device.cuh"
#include "icp_internal.h"

namespace icp {

constexpr int kWideEmpty = 0x7fffffff;   // child record without content
constexpr int kWideDone = 0x7ffffffe;    // traversal state: nothing left
constexpr int kWideStack = 48;           // stack entries per query (overflow falls back to a scan of all clusters)
constexpr int kWideThreads = 128;

// ---------------------------------------------------------------------------------------------------
// collapse (host side, one-off per static structure): binary LBVH -> L-ary tree, breadth-first numbering
// ---------------------------------------------------------------------------------------------------
void wide_build(const Bvh &b, WideBvh &w, int L, cudaStream_t s) {
    ICP_REQUIRE(L == 4 || L == 8, "wide BVH arity must be 4 or 8");
    ICP_REQUIRE(b.prim_kind == 0 && b.n >= 2, "wide BVH needs a triangle LBVH");
    const int n = b.n;
    std::vector<int2> ch(n - 1);
    std::vector<float4> nb((size_t)(2 * n - 1) * 2);
    ICP_CUDA(cudaMemcpyAsync(ch.data(), b.children.p, sizeof(int2) * (n - 1), cudaMemcpyDeviceToHost, s));
    ICP_CUDA(cudaMemcpyAsync(nb.data(), b.nodebox.p, sizeof(float4) * nb.size(), cudaMemcpyDeviceToHost, s));   // instance 0
    ICP_CUDA(cudaStreamSynchronize(s));
    // leaf range of every binary node: [first, first + count)
    std::vector<int> first(2 * n - 1), count(2 * n - 1, 0);
    for (int slot = 0; slot < n; slot++) { first[n - 1 + slot] = slot; count[n - 1 + slot] = 1; }
    {
        std::vector<int> stack{0};
        while (!stack.empty()) {
            const int v = stack.back();
            const int l = ch[v].x >= 0 ? ch[v].x : n - 1 + (~ch[v].x), r = ch[v].y >= 0 ? ch[v].y : n - 1 + (~ch[v].y);
            if (count[l] && count[r]) { count[v] = count[l] + count[r]; first[v] = std::min(first[l], first[r]); stack.pop_back(); }
            else { if (!count[l]) stack.push_back(l); if (!count[r]) stack.push_back(r); }
        }
    }
    auto area = [&](int v) {
        const float4 lo = nb[2 * (size_t)v], hi = nb[2 * (size_t)v + 1];
        const double dx = (double)hi.x - lo.x, dy = (double)hi.y - lo.y, dz = (double)hi.z - lo.z;
        return dx * dy + dy * dz + dz * dx;
    };
    std::vector<float4> out;
    std::queue<std::pair<int, int>> todo;   // (binary node, wide index)
    int n_wide = 1;
    out.resize((size_t)L * 2);
    todo.push({0, 0});
    int max_depth_items = 0;
    while (!todo.empty()) {
        const auto [bn, wi] = todo.front();
        todo.pop();
        std::vector<int> items;
        if (count[bn] <= L) items.push_back(bn);   // tiny tree: the root itself is one cluster
        else {
            const int l = ch[bn].x >= 0 ? ch[bn].x : n - 1 + (~ch[bn].x), r = ch[bn].y >= 0 ? ch[bn].y : n - 1 + (~ch[bn].y);
            items = {l, r};
            while ((int)items.size() < L) {
                int pick = -1;
                double best = -1.0;
                for (int k = 0; k < (int)items.size(); k++)
                    if (count[items[k]] > L && area(items[k]) > best) { best = area(items[k]); pick = k; }
                if (pick < 0) break;
                const int v = items[pick];
                const int l2 = ch[v].x >= 0 ? ch[v].x : n - 1 + (~ch[v].x), r2 = ch[v].y >= 0 ? ch[v].y : n - 1 + (~ch[v].y);
                items[pick] = l2;
                items.push_back(r2);
            }
        }
        max_depth_items = std::max(max_depth_items, (int)items.size());
        for (int k = 0; k < L; k++) {
            float4 a = make_float4(3e38f, 3e38f, 3e38f, -3e38f), c = make_float4(-3e38f, -3e38f, 0.f, 0.f);
            int ref = kWideEmpty;
            if (k < (int)items.size()) {
                const int v = items[k];
                const float4 lo = nb[2 * (size_t)v], hi = nb[2 * (size_t)v + 1];
                a = make_float4(lo.x, lo.y, lo.z, hi.x);
                c = make_float4(hi.y, hi.z, 0.f, 0.f);
                if (count[v] <= L) ref = ~((first[v] << 3) | (count[v] - 1));    // leaf cluster
                else {
                    ref = n_wide++;
                    out.resize((size_t)n_wide * L * 2);
                    todo.push({v, ref});
                }
            }
            memcpy(&c.z, &ref, sizeof(int));
            out[((size_t)wi * L + k) * 2] = a;
            out[((size_t)wi * L + k) * 2 + 1] = c;
        }
    }
    ICP_REQUIRE(n < (1 << 27), "too many triangles for the wide BVH's cluster references");
    w.L = L;
    w.n_nodes = n_wide;
    w.n_leaves = n;
    w.nodes.upload(out.data(), out.size(), s);
    ICP_CUDA(cudaStreamSynchronize(s));
}

// ---------------------------------------------------------------------------------------------------
// traversal
// ---------------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(kWideThreads) k_nearest_wide(int n_nodes, const float4 *__restrict__ wnodes, int n_stage,
                                                               const int *__restrict__ prim,
                                                               const double *__restrict__ prim_data, int n_leaves, int C,
                                                               long long nq, const double *__restrict__ q, int q_per_chain,
                                                               const double *__restrict__ Xq, const int *__restrict__ q_ids,
                                                               int Nq, const int *__restrict__ perm, int *__restrict__ seed_slot,
                                                               int *__restrict__ out_prim, int *__restrict__ out_feat,
                                                               double *__restrict__ out_cp, double *__restrict__ out_d2) {
    constexpr int G = kWideThreads / L;                  // queries in flight per CTA
    extern __shared__ __align__(16) float4 s_nodes[];    // [n_stage][L][2]
    __shared__ int s_ref[kWideStack][G];
    __shared__ float s_dist[kWideStack][G];
    const int tid = threadIdx.x, lane = tid & 31, grp = tid / L, gl = tid % L;
    const unsigned gmask = (L == 32 ? 0xffffffffu : ((1u << L) - 1u)) << ((lane / L) * L);   // this query's lanes in the warp
    for (int e = tid; e < n_stage * L * 2; e += kWideThreads) s_nodes[e] = __ldg(wnodes + e);
    __syncthreads();
    const long long total = nq * C;
    const long long rounds = (total + (long long)gridDim.x * G - 1) / ((long long)gridDim.x * G);
    for (long long rd = 0; rd < rounds; rd++) {
        long long g = (rd * gridDim.x + blockIdx.x) * G + grp;
        const bool live = g < total;
        if (!live) g = total - 1;
        int c = (int)(g / nq);
        long long i = g % nq;
        if (perm) { i = perm[i]; g = (long long)c * nq + i; }
        double qx, qy, qz;
        if (q_ids) {
            const double *src = Xq + ((size_t)c * Nq + q_ids[i]) * 3;
            qx = src[0]; qy = src[1]; qz = src[2];
        } else {
            const double *src = q + ((q_per_chain ? (size_t)c * nq : 0) + i) * 3;
            qx = src[0]; qy = src[1]; qz = src[2];
        }
        const float fx = (float)qx, fy = (float)qy, fz = (float)qz;
        Hit h;                                           // best among the triangles THIS lane tested
        h.d2 = INFINITY; h.x = h.y = h.z = 0.0; h.prim = 0x7fffffff; h.feat = -1; h.slot = -1;
        float best = INFINITY;                           // the query's bound (identical in its L lanes)
        int sp = 0, cur = 0;
        bool overflow = false;
        if (!live || !(qx == qx && qy == qy && qz == qz)) cur = kWideDone;
        else if (seed_slot) {
            // upper bound from the triangle that answered this query last time (every lane of the query tests it: the
            // same instructions as one lane would issue, and no exchange is needed)
            const int s0 = seed_slot[g];
            if ((unsigned)s0 < (unsigned)n_leaves) {
                leaf_test<0, false>(s0, prim, prim_data, nullptr, nullptr, qx, qy, qz, h);
                best = __double2float_ru(h.d2);
            }
        }
        auto pop = [&]() {
            int nn = kWideDone;
            while (sp > 0) {
                --sp;
                const float d = s_dist[sp][grp];
                if (d <= best) { nn = s_ref[sp][grp]; break; }
            }
            return nn;
        };
        // Every shuffle / ballot below runs in warp-uniform control flow with the full mask (xor offsets < L and the masked
        // ballots keep the data inside a query's lanes): sub-warp masks inside divergent loops would split the warp into
        // independently scheduled fragments that execute one after the other.
        while (true) {
            // phase 1: every query that holds a wide node visits it; the others idle until all hold a leaf cluster or are done
            for (;;) {
                const bool act = (unsigned)cur < (unsigned)kWideDone;
                if (!__any_sync(0xffffffffu, act)) break;
                float d = INFINITY;
                int ref = kWideEmpty;
                if (act) {
                    const float4 *rec = (cur < n_stage ? s_nodes : wnodes) + ((size_t)cur * L + gl) * 2;
                    const float4 a = rec[0], b = rec[1];
                    ref = __float_as_int(b.z);
                    d = box_d2(a.x, a.y, a.z, a.w, b.x, b.y, fx, fy, fz);
                }
                const bool hit = act && ref != kWideEmpty && d <= best;
                const unsigned hm = __ballot_sync(0xffffffffu, hit) & gmask;
                float dmin = hit ? d : INFINITY;
#pragma unroll
                for (int o = L / 2; o > 0; o >>= 1) dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
                const unsigned nm = __ballot_sync(0xffffffffu, hit && d == dmin) & gmask;
                const int near_lane = nm ? __ffs(nm) - 1 : lane;   // lane index within the warp
                const int near_ref = __shfl_sync(0xffffffffu, ref, near_lane);
                if (act) {
                    if (hm == 0) cur = pop();
                    else {
                        const unsigned others = hm & ~(1u << near_lane);
                        const int n_push = __popc(others);
                        if (sp + n_push > kWideStack) overflow = true;   // uniform over the query's lanes
                        else {
                            if (hit && lane != near_lane) {
                                const int r = sp + __popc(others & ((1u << lane) - 1u));
                                s_ref[r][grp] = ref;
                                s_dist[r][grp] = d;
                            }
                            sp += n_push;
                        }
                        cur = near_ref;
                    }
                }
                __syncwarp();
            }
            const bool at_leaf = cur != kWideDone;
            if (!__any_sync(0xffffffffu, at_leaf)) break;
            // phase 2: the exact tests of all queries of the warp together, one triangle per lane
            if (at_leaf) {
                const int code = ~cur, start = code >> 3, cnt = (code & 7) + 1;
                if (gl < cnt) leaf_test<0, false>(start + gl, prim, prim_data, nullptr, nullptr, qx, qy, qz, h);
            }
            float bl = __double2float_ru(h.d2);
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) bl = fminf(bl, __shfl_xor_sync(0xffffffffu, bl, o));
            if (at_leaf) {
                best = fminf(best, bl);
                cur = pop();
            }
            __syncwarp();
        }
        if (overflow) {
            // never seen on real meshes: a stack overflow dropped subtrees, so scan every triangle (exact, slow)
            for (int slot = gl; slot < n_leaves; slot += L) leaf_test<0, false>(slot, prim, prim_data, nullptr, nullptr, qx, qy, qz, h);
        }
        // the query's winner among its lanes: smallest distance, then lowest triangle index, then lowest lane
        double bd = h.d2;
        int bp = h.prim, bw = gl;
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int op = __shfl_xor_sync(0xffffffffu, bp, o), ow = __shfl_xor_sync(0xffffffffu, bw, o);
            if (od < bd || (od == bd && (op < bp || (op == bp && ow < bw)))) { bd = od; bp = op; bw = ow; }
        }
        if (live && gl == bw) {
            if (h.prim == 0x7fffffff) { h.d2 = NAN; h.x = h.y = h.z = NAN; h.prim = -1; }
            if (seed_slot) seed_slot[g] = h.slot;
            if (out_prim) out_prim[g] = h.prim;
            if (out_feat) out_feat[g] = h.feat;
            if (out_cp) { out_cp[3 * g] = h.x; out_cp[3 * g + 1] = h.y; out_cp[3 * g + 2] = h.z; }
            if (out_d2) out_d2[g] = h.d2;
        }
        __syncwarp();
    }
}

bool launch_nearest_wide(const NearestArgs &a, int sm_count, cudaStream_t s) {
    const WideBvh *w = a.wide;
    if (!w || w->n_nodes <= 0 || a.prim_data == nullptr || a.bvh->prim_kind != 0) return false;
    const long long total = a.nq * a.C;
    if (total <= 0) return true;
    ProfScope _ps(ST_NEAREST_STATIC, s);
    const int L = w->L, G = kWideThreads / L;
    // stage the top of the tree (breadth-first prefix, 24 KB: every level but the last of a femur-sized tree; the rest is
    // served by L1 / L2) so that five CTAs still share an SM; small launches stage the root and its children only (a CTA
    // would copy more node bytes than its few queries read)
    int n_stage = std::min(w->n_nodes, (24 * 1024) / (L * 32));
    long long blocks = (total + G - 1) / G;
    const long long persistent = (long long)sm_count * 5;
    if (blocks > persistent) blocks = persistent;
    if (total < (long long)blocks * G * 4) n_stage = std::min(n_stage, 1 + L);   // root + its children only
    const size_t smem = sizeof(float4) * 2 * (size_t)L * n_stage;
#define ICP_LAUNCH_WIDE(LL)                                                                                                   \
    do {                                                                                                                      \
        if (smem > 40 * 1024) ICP_CUDA(cudaFuncSetAttribute(k_nearest_wide<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_nearest_wide<LL><<<(unsigned)blocks, kWideThreads, smem, s>>>(w->n_nodes, w->nodes.p, n_stage, a.bvh->prim.p, a.prim_data, \
                                                                       w->n_leaves, a.C, (long long)a.nq, a.q, a.q_per_chain, a.Xq, \
                                                                       a.q_ids, a.Nq, a.perm, a.seed_slot, a.out_prim, a.out_feat,  \
                                                                       a.out_cp, a.out_d2);                                    \
    } while (0)
    if (L == 4) ICP_LAUNCH_WIDE(4); else ICP_LAUNCH_WIDE(8);
#undef ICP_LAUNCH_WIDE
    ICP_CUDA(cudaGetLastError());
    return true;
}

}  // namespace icp
