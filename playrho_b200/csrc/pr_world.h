// pr_world.h -- host-side world object and the device SoA view handed to kernels.
// Data layout in HBM (one row per id; ids are the API-visible pool indices):
//   bodies   : float4 xf | float4 pos0(x,y,a,alpha0) | float4 pos1(x,y,a,underActive) | float4 vel |
//              float4 acc | float4 mass(invM,invI,linDamp,angDamp) | float2 lc | u32 flags
//   children : radius, vertex offset/count ; pooled float2 vertices / normals
//   proxies  : body, shape, child, global child, tree id, float4 fat AABB (minx,miny,maxx,maxy)
//   contacts : u64 key | uint4 (bodyA,bodyB,childGlobalA,childGlobalB) | uint2 proxy slots |
//              uint4 api ids | float4 (friction,restitution,tangentSpeed,toi) | u32 flags |
//              manifold = 4 x float4 (ln,lp | p0,p1 | jn0,jt0,jn1,jt1 | cf0,cf1,type,count)
//   per-step : BodyConstraint (float4 vel, float4 pos), VelocityConstraint rows (144 B),
//              PositionConstraint rows (64 B), both stored in colour-sorted solve order.
#ifndef PR_WORLD_H
#define PR_WORLD_H
#include "pr_compat.h"
#include "pr_prims.cuh"
#include "../../include/playrho_b200.h"
#include <string>
#include <vector>

namespace pr {

constexpr unsigned INVALID = 0xFFFFFFFFu;
constexpr unsigned COLOR_UNSET = 0xFFu;
constexpr unsigned COLOR_OVERFLOW = 64u; // constraints that found no free colour < 64: serial tail
constexpr unsigned MAX_COLORS = 65u;
constexpr unsigned COLOR_CAP_HUBS = 24u;
constexpr unsigned LEAF_BIT = 0x80000000u;

enum { ISL_ACTIVE = 1u, ISL_VEL_DONE = 2u, ISL_POS_SOLVED = 4u };
enum { TB_RUN = 0, TB_DONE = 1, TB_HOST_FIND = 2, TB_ADJ = 3, TB_ERROR = 4, TB_BIG = 5 };
enum { ERR_NONE = 0, ERR_STACK = 1, ERR_SENSOR = 2, ERR_PAIR_OVERFLOW = 4, ERR_HASH_FULL = 8, ERR_TOI_SCRATCH = 16, ERR_FLOW_STALL = 32 };

// VelocityConstraint row (Library/include/playrho/d2/VelocityConstraint.hpp), 144 B
struct VCRow {
    float nx, ny, friction, tangentSpeed;
    unsigned bodyA, bodyB, count, contact;
    float invMA, invIA, invMB, invIB;
    float rA0x, rA0y, rB0x, rB0y;
    float jn0, jt0, nm0, tm0;
    float rA1x, rA1y, rB1x, rB1y;
    float jn1, jt1, nm1, tm1;
    float bias0, bias1, k00, k11;
    float k01, m00, m11, m01;
    unsigned island, dynA, dynB, pad;
};
// PositionConstraint row (PositionConstraint.hpp) + what GetPSM needs, 64 B
struct PCRow {
    float lnx, lny, lpx, lpy;
    float p0x, p0y, p1x, p1y;
    float lcAx, lcAy, lcBx, lcBy;
    float totalRadius; unsigned typeCount; // type | count << 8 | (rotation of body A cached) << 16 | (of body B) << 17
    float rotC, rotS;                      // UnitVec::Get(angle) of the row's STATIC body, if it has one (see solve_position_row)
};

// per-step work data of a revolute joint (RevoluteJointConf's rA, rB, mass, angularMass) + its island label
// the part of a gear joint's row that does not fit the shared 76-byte layout: the two other bodies and what the reference copies from the two
// geared joints (GearJointConf.hpp:70-83: PrismaticData / RevoluteData). type: 0 = none (std::monostate), 1 = prismatic, 2 = revolute.
// prismatic: v[0..1] = localAnchorA (on C / D), v[2..3] = localAnchorB (on A / B), v[4..5] = localAxis; revolute: v[0] = referenceAngle
struct GearExt { unsigned body_c, body_d, type_ac, type_bd; float ac[6], bd[6]; };
struct JointWork { float rAx, rAy, rBx, rBy; float m[9]; float angularMass; unsigned island; unsigned pad; float ex[8]; };

// the boxes of an LBVH node's two children, stored with the node
struct CBox { float4 l, r; };

struct Counters {
    unsigned long long toiMinKey; // (toi bits << 32 | position in m_contacts) of the soonest contact, GetSoonestContact
    unsigned toiEvents, toiScratchTop, toiAtMaxSub, toiUpdated, toiMaxDistIters, toiMaxToiIters, toiMaxRootIters;
    unsigned toiMinSepFlip, toiMaxIncBits, toiIslandsSolved, toiSumPosIters, toiSumVelIters, toiIslandsFound;
    unsigned toiUpdTouching, toiSkipTouching, toiListN, syncN, adjOverflow, toiDeferred, recolorFlag, nTailSegs;
    unsigned nQueue, nPairs, nNew, nDestroyed, nUncolored, nConstraints, nDirty, nColorWork;
    unsigned islandsFound, islandsSolved, bodiesSlept, sumVelIters, sumPosIters, maxIslandBodies;
    unsigned proxiesMoved, contactsUpdated, nAwake, nTouching, nActiveVel, nActivePos;
    unsigned minSepFlip, maxIncBits, error, pad;
    unsigned toiLogN; // TOI events logged this step (pr_get_toi_schedule)
    unsigned blkMaxBodies, blkColors; // block solver (pr_block_solve.cuh): bodies of the largest world, colours in use this step
    unsigned nStaticTouch, islStatOverflow; // (island, non-speedable body) pairs: constraints/joints seen, set overflow (see island_count_static)
    unsigned boundsMinX, boundsMinY, boundsMaxX, boundsMaxY;
    unsigned nEvContact, nEvImpulse, nEvPre, evPad; // contact events of this step (pr_events): begin/end, post-solve, pre-solve records
    unsigned lateWhy[8]; // diagnostics (PR_TOI_DEBUG): hand-backs of k_tb_late by reason, list sizes seen, events per round
    unsigned colorStart[MAX_COLORS + 1];
    // batched TOI rounds (pr_toi_phase.cuh, k_tb_*): the host enqueues several rounds without reading anything back;
    // the contact-pool bookkeeping it normally keeps (World::nOrder/nFree/poolSize) is mirrored here meanwhile
    unsigned toiHalt;       // TB_RUN, or why the remaining kernels of the batch turned into no-ops
    unsigned tbRounds, tbFound, tbAdded, tbPairs; // totals of the batch
    unsigned devOrder, devFree, devPool;          // m_contacts size, free-stack height, pool size
    unsigned tbFindSkip, tbAddOld, tbAddN, toiCandN; // toiCandN: candidates (valid TOI < 1) of the round, see k_toi_compute
    // compact lists for the single-block late rounds (k_tb_late): contacts to recompute in round toiRoundId + 1 (parity
    // ping-pong), the candidates being compacted, the listed contacts currently at maxSubSteps, "next round needs the full list"
    unsigned toiRoundId, toiReN[2], toiCand2N, toiAtMaxLive, toiForceFull, tbPad2;
    unsigned colorRoundsDev; // rounds k_color_rounds_block ran this step
    unsigned flowTicket[32]; // dataflow solver (k_flow_pass): the next block of rows of each pass
#ifdef PR_TOI_TIMING
    unsigned long long dbg[16]; // profiling build only: clock64 deltas of the phases of toi_solve_event (event 0 of each round)
    unsigned long long dbgBlk[12]; // profiling build only: clock64 deltas of the phases of k_block_solve (block 0, thread 0)
    unsigned long long dbgLate[12]; // profiling build only: clock64 deltas of the stages of a k_tb_late round (thread 0)
#endif
};

struct DV {
    // bodies
    float4 *b_xf, *b_pos0, *b_pos1, *b_vel, *b_acc, *b_mass; float2* b_lc; unsigned* b_flags;
    float4 *s_vel, *s_pos;
    unsigned *b_island, *b_root, *b_moved, *b_sync;
    pr_revolute_joint* jt; JointWork* jt_work; GearExt* jt_gear; unsigned* jt_order; unsigned* isl_jflag; // revolute joints (a15)
    unsigned long long* nc_pairs; unsigned nNcPairs;               // body pairs joined by a collideConnected=false joint
    unsigned *isl_state, *isl_velIters, *isl_posIters, *isl_inc, *isl_maxInc, *isl_sep, *isl_minSep, *isl_minUA, *isl_nbodies, *isl_active;
    unsigned long long *b_claim, *b_cmask;
    // SolveReg's maxIslandBodies counts every non-speedable body an island touches once (AabbTreeWorld.cpp:1395-1403, :1510):
    // isl_nstat[root] = distinct such bodies, deduplicated through a per-step hash set of (root, body) keys
    unsigned* isl_nstat; unsigned long long* is_keys; unsigned is_mask;
    // shapes / children / vertices
    float4* sh_mat; unsigned* sh_filter; int* sh_group;
    float* ch_radius; unsigned *ch_voff, *ch_count; float4* ch_pack; float2 *vtx, *nrm; // ch_pack: (radius, offset, count) per child, one gather
    // proxies
    unsigned *p_body, *p_shape, *p_child, *p_cg, *p_tree, *p_moved; float4* p_fat; unsigned* q_list;
    // contacts
    unsigned long long* c_key; uint4* c_ab; uint2* c_px; uint4* c_ids; float4* c_mat; unsigned* c_flags;
    float4 *m0, *m1, *m2; uint4* m3;
    unsigned *c_order, *c_free, *c_tmp, *c_tmp2;
    // TOI phase: body -> contact adjacency (m_bodyContacts) as CSR, islanded flags, event list, island scratch
    // (adjacency covers speedable bodies only, each segment has slack for contacts added during the TOI rounds)
    unsigned *adj_off, *adj_cnt, *adj_n, *adj_list, *c_pos, *b_toiIsl, *c_toiIsl, *b_toiAlpha, *b_toiStamp, *b_inR, *toi_events, *toi_islBodies,
        *toi_islContacts, *toi_list, *toi_cand, *toi_cand2, *toi_re0, *toi_re1, *c_rstamp, *sync_list;
    unsigned *bp_off, *bp_list; // body -> proxy CSR (m_bodyProxies)
    uint2* toi_log; unsigned toi_logCap; // (round, contact id) of every TOI event of the step; only while the solve order is recorded
    unsigned long long* ht_keys; unsigned ht_mask;
    // new pairs
    unsigned long long *np_key, *np_key2; uint2* np_px; unsigned *np_idx, *np_idx2;
    // constraints
    // constraint rows, structure-of-arrays (one float4 plane per 16 B of VCRow / PCRow so that a warp's access to
    // 32 consecutive rows is one coalesced 512 B transaction): see vc_load / vc_store in pr_kernels.cuh
    float4 *vc0, *vc1, *vc2, *vc3, *vc4, *vc5, *vc6, *vc7, *vc8; unsigned* vc_island;
    float4 *pc0, *pc1, *pc2, *pc3;
    unsigned* vc_flow; // dataflow solver: per row, rank | rows << 8 of body A, the same << 16 of body B (flow_word)
    unsigned *k_contact, *k_color, *k_sorted, *k_tmpKey, *k_tmpKey2, *k_tmpVal, *k_ulist;
    unsigned* c_color; // persistent colour per contact
    // contact events (pr_set_event_mask / pr_events): off unless ev_mask != 0
    pr_contact_event* ev_contact; pr_impulse_event* ev_impulse; pr_presolve_event* ev_presolve; unsigned ev_mask, ev_cap;
    // bvh
    unsigned *bv_code, *bv_code2, *bv_idx, *bv_idx2, *bv_left, *bv_right, *bv_parent, *bv_lparent, *bv_visit, *bv_leafOf;
    CBox* bv_cbox; // per internal node: the boxes of its two children (see k_bvh_refit)
    // batched independent worlds (config 5): every body carries the index of the world it lives in; the world index is
    // the high bits of the LBVH sort key, so each internal node covers a contiguous world range [bv_wr.x, bv_wr.y]
    unsigned* b_world; uint2* bv_wr; unsigned nWorlds, worldBits; // nWorlds == 0: a plain single world
    // block solver (pr_block_solve.cuh): bodies grouped by world (CSR blk_off/blk_body, b_slot = index of a body inside its
    // world), this step's first constraint row of every world, the joints of every world
    unsigned *blk_off, *blk_body, *b_slot, *blk_rowOff, *blk_jOff, *blk_joint;
    Counters* cnt;
};

template <class T> struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
};

struct World {
    pr_world_conf conf{};
    int device = 0;
    pr_stream_t stream{};
    std::string error;
    // host mirrors
    std::vector<pr_shape> shapes;       // child_offset is global
    std::vector<pr_child> children;     // vertex_offset is global
    std::vector<std::vector<unsigned>> bodyShapes;
    std::vector<pr_revolute_joint> joints; std::vector<unsigned> freeJoints; // (a destroyed joint: bodies == INVALID; its id is re-used LIFO, ObjectPool)
    std::vector<std::pair<unsigned, unsigned>> fixturesForProxies;
    // Destroy(BodyID / ShapeID), Detach, SetShape with new geometry (AabbTreeWorld.cpp:1062, :1259, :1284, :2760): ids of destroyed
    // bodies and shapes are handed out again last-freed-first (ObjectPool), and so are the DynamicTree's leaf ids -- a leaf's
    // destruction frees its parent node and then the leaf (DynamicTree.cpp:539-569), a creation takes a leaf and then a parent
    // (:517-537), so the free list pairs up and the k-th new leaf gets the k-th last destroyed LEAF id whatever the internal
    // nodes were; with the list empty a leaf takes the next fresh node, and its parent the one after it.
    struct Pools {
        std::vector<unsigned> bodyFree, shapeFree, treeFree;
        unsigned treeNextFresh = 0;
        std::vector<unsigned char> bodyDead, shapeDead;          // by id (shorter than the pool: the rest is live)
        std::vector<unsigned> pxBody, pxShape, pxTree;           // host mirror of p_body / p_shape / p_tree, by proxy slot
    } pools;
    bool needsFiltering = false; // e_needsContactFiltering: some contact is flagged for filtering (SetShape with another filter)
    unsigned nBodies = 0, nShapes = 0, nChildren = 0, nVerts = 0, nProxies = 0;
    unsigned nOrder = 0, nFree = 0, poolSize = 0;  // contacts
    unsigned nCons = 0, nColors = 0, maxTreeId = 0, debugFlags = 0;
    bool anyBodySet = false, hashDirty = true, stepComplete = true;
    bool recolorAll = true; unsigned colorsAtFull = 0;
    // Hub rule: a body touched by many constraints (a tumbler's enclosure carries ~100 squares) needs one colour per constraint,
    // i.e. one dependent launch per constraint and iteration. When a world needs more than 32 colours and the constraints beyond
    // colour COLOR_CAP_HUBS are few per island, colours stop at the cap and the rest go to the per-island serial tail (one thread
    // per island, list order: still a valid Gauss-Seidel order) -- 65 colours + tail become 24 + tail for an arena of tumblers.
    // Measured on the 8 192-tumbler arena: 29.8 ms/step with the rule, 29.2 without (the serial tail's dependent row loads cost what
    // the 40 saved launches per iteration cost): opt-in (pr_set_debug bit 4), bit-exact, under test.
    unsigned colorCap = 64u;
    bool subStepping = false, adjValid = false, bvhValid = false, bpDirty = true, toiAlphaDirty = true;
    unsigned bvhFullRefits = 0; // steps since the last rebuild that only re-fitted the boxes over the same topology
    unsigned bvhLeaves = 0, bvhRefits = 0, hashTombs = 0, toiListBound = 0, toiRoundsHint = 1, toiGroupsHint = 1;
    const unsigned* bvhSortedIdx = nullptr;
    std::vector<std::vector<unsigned>> bodyProxies; // m_bodyProxies
    std::vector<std::vector<unsigned>> bodyJointList; // m_bodyJoints: a body's joints in the order they were added to it (Destroy(BodyID) frees them in it)
    unsigned toiMode = 1; // PR_TOI_PARALLEL (default) or PR_TOI_SERIAL_EXACT (reference order)
    // some body WITH shapes is (or may be: SetBody, load) impenetrable -- static, kinematic or bullet: only then can a contact be a
    // TOI candidate (UpdateContactTOIs' filter :1696). An arena of tumblers has none: its TOI phase is skipped outright.
    bool toiPossible = false;
    float inv_dt0 = 0;
    // capacities
    unsigned capBodies = 0, capShapes = 0, capChildren = 0, capVerts = 0, capProxies = 0, capContacts = 0,
             capPairs = 0, capCons = 0, capHash = 0, capAdj = 0, capJoints = 0;
    std::vector<unsigned> jointLevelOff; // joints that share a movable body are solved level by level (refresh_joint_tables): offsets into d.jt_order
    unsigned nJoints = 0; bool jointsSerial = false, jointTablesDirty = true; // (tables: jointsSerial, the collideConnected == false pair list)
    std::vector<unsigned> bodyWorld; // batch arena only: world index of each body
    DV d{};
    RadixScratch rs;
    unsigned* scanScratch = nullptr; size_t scanScratchCap = 0;
    Counters hostCnt{};
    unsigned long long cntSerial = 0; bool cntSerialValid = false; // launch serial at the last device->host read of the counters
    static constexpr unsigned PINNED_SLOTS = 64;
    Counters* pinnedCnt = nullptr; unsigned pinnedNext = 1;
    pr_step_profile profile{};
    std::vector<unsigned> solveOrderIds, solveOrderColors;
    std::vector<uint2> toiSchedule; unsigned toiScheduleN = 0; // last step's TOI events by round (debug bit 0)
    std::vector<void*> allocs;
    const unsigned* blkSortedKey = nullptr; // this step's sorted (world << 7 | colour) keys
    bool tmaSmemSet = false;
    bool contactsClean = false; // no contact is flagged for updating and no body moved since the last Step ended (step_impl)
    bool blocksDirty = true, blkJointsDirty = true; unsigned blkMaxBodies = 0, blkWorldCap = 0, blkSmemSet = 0; // block solver
    unsigned isCap = 0, isHint = 0; // island-static pair set: capacity (power of two), pairs seen last step
    unsigned* gridBar = nullptr; unsigned fusedGrid = 0; bool fusedProbed = false; // persistent solver kernel (pr_solve_fused.cuh)
    void* timer = nullptr; // StageTimer (pr_world.cu): per-stage CUDA events of this world's device
    void* stageBuf[2] = {nullptr, nullptr}; size_t stageCap[2] = {0, 0}; // persistent device staging for row upload/download
};

void count_launch();
unsigned launches_since_reset();
void reset_launch_count();

} // namespace pr
#endif
