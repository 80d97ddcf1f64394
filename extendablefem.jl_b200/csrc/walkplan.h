// walkplan.h -- host-side planning of "walk programs" for the template kernel of 3D P2 columns.
//
// The template kernel (fastplan.cuh: tp_gather_kernel) adds every cell-local entry to its accumulator with a shared-memory
// read-modify-write because accumulator positions are template DATA.  For the P2 tetrahedral Laplacian most of those
// round trips can be avoided by visiting the cells of a column in an order in which consecutive cells share rows, and
// keeping the shared rows in REGISTERS with fixed roles:
//
//   edge-dof column (edge pq): the cells around the edge form a ring (p, q, v_i, v_i+1).  Walking the ring, the rows of the
//     vertex shared with the next cell (v, edge pv, edge qv) are carried in three registers; the rows of the vertex shared
//     with the previous cell are completed by the carry and stored; the ring edge v_i v_i+1 belongs to this cell only.
//     Every row becomes store-only (the rows p, q, pq accumulate in registers for the whole column as before).
//   vertex-dof column (vertex p): the cells around p are the triangles (x, y, z) of its link.  Walking from triangle to
//     triangle across shared link edges, the two vertices of the crossed edge stay in a register window (vertex row +
//     spoke row px each), the crossed edge's row is carried, a vertex's rows are flushed when it leaves the window and
//     re-loaded if the walk comes back to it.
//
// The local column is evaluated in ROLE order: the kernel loads the entries D_ab of the barycentric Gram matrix through
// plane indices stored per round (which local pair plays which role), so there is ONE code path per column class instead
// of one per local index.  The planner below turns a template (the rounds of one column shape) into walk records, and
// VERIFIES them by simulating the kernel symbolically: every (round, local row) contribution must arrive exactly once at its
// position.  A template that cannot be verified keeps the read-modify-write kernel (for the whole plan).
#pragma once
#include <algorithm>
#include <array>
#include <cstring>
#include <map>
#include <set>
#include <vector>

namespace extfem {

constexpr int TW_RW = 12;  // 32-bit words of one walk record (== TP_TW: walk records share the templates' constant bank)
constexpr int TW_POS_BYTES = 33 * 8;   // bytes between accumulator positions (TP_LD * 8)
// flags (word 1, low 16 bits)
enum : unsigned {
    TWF_WIN = 1u << 0,    // vertex: PL / PS rows are in the register window          | edge: IN rows are the carry registers
    TWF_LD0 = 1u << 1,    // vertex: PL rows start from the stored partial sums        | edge: IN rows start from stored sums
    TWF_LD1 = 1u << 2,    // vertex: PS rows start from the stored partial sums        | edge: OUT rows start from stored sums
    TWF_LD2 = 1u << 3,    // vertex: NW rows start from the stored partial sums        | edge: ring edge starts from stored sum
    TWF_CIN = 1u << 4,    // vertex: in-edge (PL,PS) adds the carried edge register
    TWF_LDEI = 1u << 5,   // vertex: in-edge starts from the stored partial sum
    TWF_LDEO = 1u << 6,   // vertex: out-edge (PS,NW) starts from the stored partial sum
    TWF_CO = 1u << 7,     // vertex: out-edge is carried to the next round             | edge: OUT rows are carried
    TWF_LDET = 1u << 8,   // vertex: third edge (PL,NW) starts from the stored partial sum
    TWF_FL1 = 1u << 9,    // vertex: PS rows are stored after this round
    TWF_FL2 = 1u << 10,   // vertex: NW rows are stored after this round
    TWF_SWAP = 1u << 11,  // vertex: next window = (NW, PS) instead of (PS, NW)
    TWF_EDGE = 1u << 15   // the column is an edge-dof column
};

struct WalkTemplateIn {
    int m = 0, L = 0;
    std::vector<int> celloff;              // [m] transposed cell offset (template word 0)
    std::vector<int> kl;                   // [m] local index of the column dof
    std::vector<int> orient;               // [m] P2 edge columns: local edge vertices swapped relative to round 0
    std::vector<std::array<int, 10>> pos;  // [m][10] accumulator position of local row t
};

inline int tw_edge_index(int x, int y)   // local P2 dof of the tet edge (x, y)
{
    if (x > y) std::swap(x, y);
    static const int idx[4][4] = {{-1, 4, 5, 6}, {4, -1, 7, 8}, {5, 7, -1, 9}, {6, 8, 9, -1}};
    return idx[x][y];
}
inline int tw_pair_plane(int x, int y)   // plane of D_xy in the geometry record (fp_pair_index<3>)
{
    if (x > y) std::swap(x, y);
    return x * (7 - x) / 2 + (y - x - 1);
}
inline void tw_edge_vertices(int e, int &a, int &b)   // local edge e = 4..9 -> (a, b), fastpath.cuh: fp_edge_a / fp_edge_b
{
    const int k = e - 4;
    a = k < 3 ? 0 : (k < 5 ? 1 : 2);
    b = k < 3 ? k + 1 : (k < 5 ? k - 1 : 3);
}

struct WalkRec {
    int celloff = 0;
    unsigned flags = 0;
    int planes[5] = {0, 0, 0, 0, 0};
    int pos[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int cpos[3] = {0, 0, 0};   // positions of the register rows of the column (first record only)
    // bookkeeping of the verifier: the (round, local row) contribution behind every value slot
    int src_round = 0;
    int t_common[3] = {-1, -1, -1};
    int t_rows[10] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
};

// record layout (what the kernel reads; everything is pre-multiplied so that the hot loop does no index arithmetic):
//   w0        flags
//   w1..w5    element offsets of the Gram-matrix entries of the round relative to the column's base cell:
//             plane * Npad + transposed cell offset (roles: edge ab ac ad bc bd | vertex px py pz)
//   w6..w10   accumulator byte offsets (position * TP_LD * 8) as 16-bit halves p0|p1, p2|p3, p4|p5, p6|p7, p8|c0
//   w11       c1 | c2   (c0..c2: register rows of the column, first record only)
inline void tw_pack(const WalkRec &R, bool first, long long Npad, unsigned *w)
{
    memset(w, 0, TW_RW * 4);
    w[0] = R.flags;
    for (int i = 0; i < 5; ++i) w[1 + i] = (unsigned)(int)((long long)R.planes[i] * Npad + R.celloff);
    unsigned short h[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 9; ++i) h[i] = (unsigned short)(R.pos[i] * TW_POS_BYTES);
    if (first) for (int i = 0; i < 3; ++i) h[9 + i] = (unsigned short)(R.cpos[i] * TW_POS_BYTES);
    for (int i = 0; i < 6; ++i) w[6 + i] = (unsigned)h[2 * i] | ((unsigned)h[2 * i + 1] << 16);
}

// ---- edge-dof column: ring walk ------------------------------------------------------------------------------------------
inline bool tw_plan_edge(const WalkTemplateIn &T, std::vector<WalkRec> &out)
{
    const int m = T.m;
    struct Cell { int pl, ql, c, d; };   // local vertices: p side, q side, the two ring vertices
    std::vector<Cell> C(m);
    for (int r = 0; r < m; ++r) {
        if (T.kl[r] < 4) return false;
        int a, b;
        tw_edge_vertices(T.kl[r], a, b);
        C[r].pl = T.orient[r] ? b : a;
        C[r].ql = T.orient[r] ? a : b;
        int o[2], n = 0;
        for (int v = 0; v < 4; ++v) if (v != a && v != b) o[n++] = v;
        C[r].c = o[0]; C[r].d = o[1];
        // the register rows must sit at the same positions in every round
        if (T.pos[r][T.kl[r]] != T.pos[0][T.kl[0]] || T.pos[r][C[r].pl] != T.pos[0][C[0].pl] || T.pos[r][C[r].ql] != T.pos[0][C[0].ql]) return false;
    }
    auto vpos = [&](int r, int v) { return T.pos[r][v]; };
    // ring order: every connected piece is walked from an open end if it has one
    std::vector<int> order, IN(m, -1), OUT(m, -1);
    std::vector<char> used(m, 0);
    auto other = [&](int r, int x) { return x == C[r].c ? C[r].d : C[r].c; };
    while ((int)order.size() < m) {
        int start = -1;
        for (int r = 0; r < m && start < 0; ++r) {
            if (used[r]) continue;
            int nb = 0;
            for (int s = 0; s < m; ++s) {
                if (s == r || used[s]) continue;
                for (int x : {C[r].c, C[r].d})
                    for (int y : {C[s].c, C[s].d}) if (vpos(r, x) == vpos(s, y)) ++nb;
            }
            if (nb <= 1) start = r;
        }
        if (start < 0) for (int r = 0; r < m; ++r) if (!used[r]) { start = r; break; }
        int cur = start;
        used[cur] = 1; order.push_back(cur);
        for (;;) {
            int nxt = -1, vo = -1, vi = -1;
            for (int s = 0; s < m && nxt < 0; ++s) {
                if (used[s]) continue;
                for (int x : {C[cur].c, C[cur].d}) {
                    if (x == IN[cur] || nxt >= 0) continue;
                    for (int y : {C[s].c, C[s].d})
                        if (nxt < 0 && vpos(cur, x) == vpos(s, y)) { nxt = s; vo = x; vi = y; }
                }
            }
            if (nxt < 0) {
                if (IN[cur] < 0) IN[cur] = C[cur].c;
                OUT[cur] = other(cur, IN[cur]);
                break;
            }
            OUT[cur] = vo;
            if (IN[cur] < 0) IN[cur] = other(cur, vo);
            IN[nxt] = vi;
            cur = nxt; used[cur] = 1; order.push_back(cur);
        }
    }
    std::set<int> touched;
    out.assign(m, WalkRec());
    for (int i = 0; i < m; ++i) {
        const int r = order[i];
        const Cell &c = C[r];
        WalkRec &R = out[i];
        R.celloff = T.celloff[r];
        R.src_round = r;
        R.flags = TWF_EDGE;
        const int in = IN[r], ou = OUT[r];
        const bool cin = i > 0 && (out[i - 1].flags & TWF_CO) != 0;
        const int rows_in[3] = {in, tw_edge_index(c.pl, in), tw_edge_index(c.ql, in)};
        const int rows_out[3] = {ou, tw_edge_index(c.pl, ou), tw_edge_index(c.ql, ou)};
        const int ring = tw_edge_index(in, ou);
        for (int k = 0; k < 3; ++k) { R.pos[k] = T.pos[r][rows_in[k]]; R.pos[3 + k] = T.pos[r][rows_out[k]]; }
        R.pos[6] = T.pos[r][ring];
        if (cin) R.flags |= TWF_WIN;
        else if (touched.count(R.pos[0])) R.flags |= TWF_LD0;
        // carry out when the next ordered round takes this OUT vertex as its IN vertex
        bool co = false;
        if (i + 1 < m) { const int s = order[i + 1]; co = vpos(s, IN[s]) == vpos(r, ou); }
        if (touched.count(R.pos[3])) R.flags |= TWF_LD1;
        if (co && !(R.flags & TWF_LD1)) R.flags |= TWF_CO;      // a re-entered vertex is stored, not carried (keeps the load simple)
        else co = false;
        if (touched.count(R.pos[6])) R.flags |= TWF_LD2;
        for (int k = 0; k < 3; ++k) touched.insert(R.pos[k]);
        if (!co) for (int k = 0; k < 3; ++k) touched.insert(R.pos[3 + k]);
        touched.insert(R.pos[6]);
        R.planes[0] = tw_pair_plane(c.pl, c.ql); R.planes[1] = tw_pair_plane(c.pl, in); R.planes[2] = tw_pair_plane(c.pl, ou);
        R.planes[3] = tw_pair_plane(c.ql, in); R.planes[4] = tw_pair_plane(c.ql, ou);
        R.cpos[0] = T.pos[r][T.kl[r]]; R.cpos[1] = T.pos[r][c.pl]; R.cpos[2] = T.pos[r][c.ql];
        R.t_common[0] = T.kl[r]; R.t_common[1] = c.pl; R.t_common[2] = c.ql;
        for (int k = 0; k < 3; ++k) { R.t_rows[k] = rows_in[k]; R.t_rows[3 + k] = rows_out[k]; }
        R.t_rows[6] = ring;
    }
    // a round whose IN rows are the carry must follow a round that carried out (checked above through TWF_CO of i-1)
    return true;
}

// ---- vertex-dof column: walk over the link triangles ------------------------------------------------------------------------
inline bool tw_plan_vertex(const WalkTemplateIn &T, std::vector<WalkRec> &out)
{
    const int m = T.m;
    struct Tri { int v[3]; };   // local vertices other than the column vertex
    std::vector<Tri> C(m);
    for (int r = 0; r < m; ++r) {
        if (T.kl[r] >= 4) return false;
        int n = 0;
        for (int v = 0; v < 4; ++v) if (v != T.kl[r]) C[r].v[n++] = v;
        if (T.pos[r][T.kl[r]] != T.pos[0][T.kl[0]]) return false;
    }
    auto gid = [&](int r, int v) { return T.pos[r][v]; };   // position of the vertex row identifies the global vertex
    auto has = [&](int r, int g) { for (int k = 0; k < 3; ++k) if (gid(r, C[r].v[k]) == g) return C[r].v[k]; return -1; };
    auto nshared = [&](int r, int s) { int n = 0; for (int k = 0; k < 3; ++k) if (has(s, gid(r, C[r].v[k])) >= 0) ++n; return n; };
    std::vector<char> used(m, 0);
    auto remaining = [&](int g) { int n = 0; for (int s = 0; s < m; ++s) if (!used[s] && has(s, g) >= 0) ++n; return n; };
    std::vector<int> order;
    std::vector<char> brk(m, 0);    // no window between ordered round i-1 and i
    int cur = 0;
    used[0] = 1; order.push_back(0); brk[0] = 1;
    int in_a = -1, in_b = -1;       // global ids of the edge shared with the previous round (-1: none)
    while ((int)order.size() < m) {
        int best = -1, bestscore = 1 << 30;
        for (int s = 0; s < m; ++s) {
            if (used[s] || nshared(cur, s) != 2) continue;
            // the vertex of `cur` that is not in s leaves the window: it should not be needed again
            int leave = -1;
            for (int k = 0; k < 3; ++k) if (has(s, gid(cur, C[cur].v[k])) < 0) leave = gid(cur, C[cur].v[k]);
            // crossing back over the in-edge is impossible (s would be the previous round); prefer leaving vertices that are done
            int score = remaining(leave) * 16;
            for (int k = 0; k < 3; ++k) if (has(cur, gid(s, C[s].v[k])) < 0) score += remaining(gid(s, C[s].v[k]));
            if (in_a >= 0 && has(s, in_a) >= 0 && has(s, in_b) >= 0) continue;
            if (score < bestscore) { bestscore = score; best = s; }
        }
        const int i = (int)order.size();
        if (best < 0) {
            for (int s = 0; s < m; ++s) if (!used[s]) { best = s; break; }
            brk[i] = 1; in_a = in_b = -1;
        } else {
            int k2 = 0, e[2] = {-1, -1};
            for (int k = 0; k < 3; ++k) if (has(best, gid(cur, C[cur].v[k])) >= 0 && k2 < 2) e[k2++] = gid(cur, C[cur].v[k]);
            in_a = e[0]; in_b = e[1];
        }
        used[best] = 1; order.push_back(best); cur = best;
    }
    // roles per ordered round
    out.assign(m, WalkRec());
    std::set<int> touched;
    int win0 = -1, win1 = -1;   // global ids in the register window before the round (slot 0 = PL, slot 1 = PS)
    bool carry = false;
    for (int i = 0; i < m; ++i) {
        const int r = order[i];
        WalkRec &R = out[i];
        R.celloff = T.celloff[r];
        R.src_round = r;
        const bool win = !brk[i];
        const bool nextwin = i + 1 < m && !brk[i + 1];
        int PL = -1, PS = -1, NW = -1;   // local vertices
        if (win) {
            PL = has(r, win0); PS = has(r, win1);
            if (PL < 0 || PS < 0) return false;
            for (int k = 0; k < 3; ++k) if (C[r].v[k] != PL && C[r].v[k] != PS) NW = C[r].v[k];
        } else {
            // fresh start: the vertex that is not shared with the next round leaves first
            if (nextwin) {
                const int s = order[i + 1];
                for (int k = 0; k < 3; ++k) if (has(s, gid(r, C[r].v[k])) < 0) PL = C[r].v[k];
            }
            if (PL < 0) PL = C[r].v[0];
            for (int k = 0; k < 3; ++k) if (C[r].v[k] != PL) { if (PS < 0) PS = C[r].v[k]; else NW = C[r].v[k]; }
        }
        if (NW < 0) return false;
        bool swap = false;
        if (nextwin) {
            // the next round keeps two of (PS, NW) ... its PL is the one of them that is not in the round after next, but
            // window slot 0 must simply hold the next round's PL: decide the next round's PL now
            const int s = order[i + 1];
            if (has(s, gid(r, PS)) < 0 || has(s, gid(r, NW)) < 0) {
                // the next round shares (PL, x): only possible at a fresh start, where PL was chosen as the unshared vertex
                return false;
            }
            // next PL: the one of (PS, NW) that is not in the round after next (if that round continues the window)
            int nextPL = gid(r, PS);
            if (i + 2 < m && !brk[i + 2]) {
                const int s2 = order[i + 2];
                if (has(s2, gid(r, PS)) >= 0 && has(s2, gid(r, NW)) < 0) nextPL = gid(r, NW);
                else if (has(s2, gid(r, PS)) >= 0 && has(s2, gid(r, NW)) >= 0) return false;   // would cross back
            }
            swap = nextPL == gid(r, NW);
        }
        R.flags = 0;
        const int t_PLv = PL, t_PLs = tw_edge_index(T.kl[r], PL), t_PSv = PS, t_PSs = tw_edge_index(T.kl[r], PS), t_NWv = NW,
                  t_NWs = tw_edge_index(T.kl[r], NW), t_Ei = tw_edge_index(PL, PS), t_Eo = tw_edge_index(PS, NW), t_Et = tw_edge_index(PL, NW);
        const int tt[9] = {t_PLv, t_PLs, t_PSv, t_PSs, t_NWv, t_NWs, t_Ei, t_Eo, t_Et};
        for (int k = 0; k < 9; ++k) { R.pos[k] = T.pos[r][tt[k]]; R.t_rows[k] = tt[k]; }
        if (win) R.flags |= TWF_WIN;
        else {
            if (touched.count(R.pos[0])) R.flags |= TWF_LD0;
            if (touched.count(R.pos[2])) R.flags |= TWF_LD1;
        }
        if (touched.count(R.pos[4])) R.flags |= TWF_LD2;
        if (win && carry) R.flags |= TWF_CIN;
        else if (touched.count(R.pos[6])) R.flags |= TWF_LDEI;
        if (touched.count(R.pos[7])) R.flags |= TWF_LDEO;
        if (touched.count(R.pos[8])) R.flags |= TWF_LDET;
        const bool co = nextwin;
        if (co) R.flags |= TWF_CO;
        if (!nextwin) R.flags |= TWF_FL1 | TWF_FL2;
        if (swap) R.flags |= TWF_SWAP;
        touched.insert(R.pos[0]); touched.insert(R.pos[1]);
        if (!nextwin) { touched.insert(R.pos[2]); touched.insert(R.pos[3]); touched.insert(R.pos[4]); touched.insert(R.pos[5]); }
        touched.insert(R.pos[6]);
        if (!co) touched.insert(R.pos[7]);
        touched.insert(R.pos[8]);
        R.planes[0] = tw_pair_plane(T.kl[r], PL); R.planes[1] = tw_pair_plane(T.kl[r], PS); R.planes[2] = tw_pair_plane(T.kl[r], NW);
        R.cpos[0] = T.pos[r][T.kl[r]];
        R.t_common[0] = T.kl[r];
        carry = co;
        if (nextwin) { win0 = swap ? gid(r, NW) : gid(r, PS); win1 = swap ? gid(r, PS) : gid(r, NW); }
        else { win0 = win1 = -1; }
    }
    return true;
}

// ---- verification: symbolic execution of the kernel's round semantics ---------------------------------------------------
// values are multisets of contribution ids (round * 16 + local row); at the end every position must hold exactly the
// contributions that the template's (round, row) -> position map sends there
inline bool tw_verify(const WalkTemplateIn &T, const std::vector<WalkRec> &W, bool first_mode)
{
    typedef std::multiset<int> Val;
    std::map<int, Val> acc;       // stored values; missing == never stored
    std::set<int> preset;         // !first_mode: positions hold an initial value that every row must pick up exactly once
    const int INIT = -1;
    if (!first_mode) for (int p = 0; p < T.L; ++p) acc[p] = Val{INIT};
    auto load = [&](int p, bool flag) -> Val {
        const bool doload = flag || !first_mode;
        if (!doload) return Val();
        auto it = acc.find(p);
        if (it == acc.end()) return Val{-999};   // loads an uninitialised accumulator: invalid program
        return it->second;
    };
    auto add = [](Val a, const Val &b) { a.insert(b.begin(), b.end()); return a; };
    auto contrib = [](int round, int t) { return Val{round * 16 + t}; };
    Val creg[3], win[4], carry1, carry3[3];
    const bool edge = (W[0].flags & TWF_EDGE) != 0;
    for (size_t i = 0; i < W.size(); ++i) {
        const WalkRec &R = W[i];
        const int r = R.src_round;
        const unsigned f = R.flags;
        if (((f & TWF_EDGE) != 0) != edge) return false;
        if (edge) {
            for (int k = 0; k < 3; ++k) creg[k] = add(creg[k], contrib(r, R.t_common[k]));
            Val in[3], ou[3];
            for (int k = 0; k < 3; ++k) {
                in[k] = (f & TWF_WIN) ? carry3[k] : load(R.pos[k], (f & TWF_LD0) != 0);
                in[k] = add(in[k], contrib(r, R.t_rows[k]));
                acc[R.pos[k]] = in[k];
                ou[k] = add(load(R.pos[3 + k], (f & TWF_LD1) != 0), contrib(r, R.t_rows[3 + k]));
                if (f & TWF_CO) carry3[k] = ou[k]; else acc[R.pos[3 + k]] = ou[k];
            }
            acc[R.pos[6]] = add(load(R.pos[6], (f & TWF_LD2) != 0), contrib(r, R.t_rows[6]));
        } else {
            creg[0] = add(creg[0], contrib(r, R.t_common[0]));
            if (!(f & TWF_WIN)) {
                win[0] = load(R.pos[0], (f & TWF_LD0) != 0); win[1] = load(R.pos[1], (f & TWF_LD0) != 0);
                win[2] = load(R.pos[2], (f & TWF_LD1) != 0); win[3] = load(R.pos[3], (f & TWF_LD1) != 0);
            }
            Val pl0 = add(win[0], contrib(r, R.t_rows[0])), pl1 = add(win[1], contrib(r, R.t_rows[1]));
            acc[R.pos[0]] = pl0; acc[R.pos[1]] = pl1;
            Val ps0 = add(win[2], contrib(r, R.t_rows[2])), ps1 = add(win[3], contrib(r, R.t_rows[3]));
            Val nw0 = add(load(R.pos[4], (f & TWF_LD2) != 0), contrib(r, R.t_rows[4]));
            Val nw1 = add(load(R.pos[5], (f & TWF_LD2) != 0), contrib(r, R.t_rows[5]));
            Val ei = (f & TWF_CIN) ? carry1 : load(R.pos[6], (f & TWF_LDEI) != 0);
            acc[R.pos[6]] = add(ei, contrib(r, R.t_rows[6]));
            Val eo = add(load(R.pos[7], (f & TWF_LDEO) != 0), contrib(r, R.t_rows[7]));
            if (f & TWF_CO) carry1 = eo; else acc[R.pos[7]] = eo;
            acc[R.pos[8]] = add(load(R.pos[8], (f & TWF_LDET) != 0), contrib(r, R.t_rows[8]));
            if (f & TWF_FL1) { acc[R.pos[2]] = ps0; acc[R.pos[3]] = ps1; }
            if (f & TWF_FL2) { acc[R.pos[4]] = nw0; acc[R.pos[5]] = nw1; }
            if (f & TWF_SWAP) { win[0] = nw0; win[1] = nw1; win[2] = ps0; win[3] = ps1; }
            else { win[0] = ps0; win[1] = ps1; win[2] = nw0; win[3] = nw1; }
        }
    }
    const int nc = edge ? 3 : 1;
    for (int k = 0; k < nc; ++k) acc[W[0].cpos[k]] = add(load(W[0].cpos[k], false), creg[k]);
    // expected
    std::map<int, Val> want;
    if (!first_mode) for (int p = 0; p < T.L; ++p) want[p] = Val{INIT};
    for (int r = 0; r < T.m; ++r)
        for (int t = 0; t < 10; ++t) want[T.pos[r][t]].insert(r * 16 + t);
    if (first_mode) {
        // every position of the column must have been stored (first-touch stores replace the zeroing pass)
        for (int p = 0; p < T.L; ++p) if (!want.count(p)) return false;
    }
    return acc == want;
}

// walk records of one template ([m][TW_RW] words); false: keep the read-modify-write kernel
inline bool tw_plan_template(const WalkTemplateIn &T, long long Npad, std::vector<unsigned> &words)
{
    if (T.m < 1 || T.L * TW_POS_BYTES > 65535 || 6 * Npad + (1ll << 28) >= (1ll << 31)) return false;
    std::vector<WalkRec> W;
    const bool edge = T.kl[0] >= 4;
    if (!(edge ? tw_plan_edge(T, W) : tw_plan_vertex(T, W))) return false;
    if (!tw_verify(T, W, true) || !tw_verify(T, W, false)) return false;
    words.assign((size_t)T.m * TW_RW, 0u);
    for (int i = 0; i < T.m; ++i) tw_pack(W[i], i == 0, Npad, &words[(size_t)i * TW_RW]);
    return true;
}

} // namespace extfem
