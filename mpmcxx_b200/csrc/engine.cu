// engine.cu — host side of the B200 energy engine and its C-ABI (include/mpmc_b200.h).
//
// One engine = the device-resident image of one reference `System` (or of P_local path-integral bead systems that
// share a topology).  The host keeps a shadow of the flat site table (the reference's linked lists stay the source of
// truth in the caller); the device holds
//     posq  double4[n_beads][cap]   x, y, z, q                (32 B/site, the only array a move rewrites)
//     lj    double2[cap]            sqrt(eps) (0 if LJ-null), sigma/2
//     alpha double [cap], mass double[cap], meta int[cap]  (molecule index | frozen<<31)
// plus O(N) work arrays for the polarization solve.  Nothing O(N^2) is ever stored (the reference keeps 200 B per
// pair and a 3N x 3N matrix).  energy() = a fixed sequence of kernels on one stream; every reduction has a fixed
// shape, so the same configuration always gives the same bits.
#include "../../include/mpmc_b200.h"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "kernels_gs.cuh"
#include "kernels_pair.cuh"
#include "kernels_pair2.cuh"
#include "kernels_polar.cuh"
#include "kernels_polar2.cuh"
#include "kernels_recip.cuh"
#include "kernels_pi.cuh"

using namespace mpmc;

static thread_local std::string g_err;

#define CK(call)                                                                                           \
	do {                                                                                                   \
		cudaError_t _e = (call);                                                                           \
		if (_e != cudaSuccess) {                                                                           \
			char _b[512];                                                                                  \
			snprintf(_b, sizeof _b, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
			g_err = _b;                                                                                    \
			return MPMC_ERR_CUDA;                                                                          \
		}                                                                                                  \
	} while (0)

#define FAIL(code, ...)                        \
	do {                                       \
		char _b[512];                          \
		snprintf(_b, sizeof _b, __VA_ARGS__);  \
		g_err = _b;                            \
		return (code);                         \
	} while (0)

namespace {

// NCCL is reached through dlopen so that single-GPU users need no NCCL at all; only the bead-sharded path-integral calls use it.
typedef struct ncclComm *nccl_comm_t;
struct NcclId { char internal[128]; };          // ncclUniqueId (nccl.h), passed by value
struct NcclApi {
	void *lib = nullptr;
	int (*GetUniqueId)(void *) = nullptr;
	int (*CommInitRank)(nccl_comm_t *, int, NcclId, int) = nullptr;
	int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
	int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
	int (*CommDestroy)(nccl_comm_t) = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values (nccl.h)

int load_nccl() {
	if (g_nccl.lib) return MPMC_OK;
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!h) { g_err = std::string("cannot load libnccl: ") + dlerror(); return MPMC_ERR_UNSUPPORTED; }
	g_nccl.GetUniqueId = (int (*)(void *))dlsym(h, "ncclGetUniqueId");
	g_nccl.CommInitRank = (int (*)(nccl_comm_t *, int, NcclId, int))dlsym(h, "ncclCommInitRank");
	g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
	g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllGather");
	g_nccl.CommDestroy = (int (*)(nccl_comm_t))dlsym(h, "ncclCommDestroy");
	g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
	if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.CommDestroy) {
		g_err = "libnccl is missing required symbols"; return MPMC_ERR_UNSUPPORTED;
	}
	g_nccl.lib = h;
	return MPMC_OK;
}
#define NK(call)                                                                                              \
	do {                                                                                                      \
		int _r = (call);                                                                                      \
		if (_r != 0) {                                                                                        \
			g_err = std::string("NCCL: ") + #call + " -> " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "error"); \
			return MPMC_ERR_CUDA;                                                                             \
		}                                                                                                     \
	} while (0)

template <class T> struct DevBuf {
	T *p = nullptr;
	size_t cap = 0;
	int ensure(size_t count) {
		if (count <= cap) return MPMC_OK;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = std::max<size_t>(count, 16);
		if (cudaMalloc(&p, want * sizeof(T)) != cudaSuccess) { cudaGetLastError(); g_err = "cudaMalloc failed"; return MPMC_ERR_ALLOC; }
		cap = want;
		return MPMC_OK;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// per-energy() scalars written by the kernels and copied back once: kResStride doubles per bead
// (rd_pair, es_real, es_intra, n_in, es_recip, sum mu.E_s, sum mu.dE_ind, sum rrms)
inline size_t res_len(int B) { return (size_t)kResStride * B; }

} // namespace

struct mpmc_engine {
	mpmc_config cfg;
	int B = 1, n = 0, cap = 0, dev = 0, num_sms = 0;
	cudaStream_t stream = nullptr;
	long long launches = 0;
	// cell
	CellDev cell;
	bool ortho = false;
	std::vector<KVec> kvec;
	// host shadow (list order)
	std::vector<double> h_pos;   // [B][n][3]
	std::vector<double> h_q, h_alpha, h_eps, h_sigma, h_mass;
	std::vector<int> h_mol, h_frozen;
	// derived host state
	std::vector<int> plist, mobile_q, frozen_q, mol_start;
	std::vector<unsigned char> mol_mobile;
	double lrc_pair = 0, lrc_self = 0, es_self = 0, n_pair_evals = 0;
	bool topo_dirty = true, frozen_sk_dirty = true, cell_tables_dirty = true;
	// device
	DevBuf<double4> d_posq;
	DevBuf<double2> d_lj;
	DevBuf<double> d_alpha, d_mass;
	DevBuf<int> d_meta, d_plist, d_mobile_q, d_frozen_q, d_mol_start, d_order, d_flags;
	// second-generation pair sweep (kernels_pair2.cuh)
	std::vector<PairSeg> segs;
	PairParams pp;
	RadialTable erf_tab;
	DevBuf<PairSeg> d_segs;
	DevBuf<int> d_pmeta, d_item_ctr, d_perm;
	DevBuf<PairItem> d_items;
	int pair_ctr_start = 0;
	DevBuf<double2> d_slj;
	DevBuf<double4> d_spq, d_stage;
	DevBuf<int> d_iperm;             // site -> position in the class-sorted table (k_scatter_sites keeps d_spq current)
	bool perm_identity = true, spq_valid = false;
	// mobile structure factor: the chunk partials stay valid between evaluations; only chunks holding a moved site are recomputed
	bool sk_valid = false;
	std::vector<int> sk_dirty;
	DevBuf<int> d_sk_dirty, d_pi_done;
	DevBuf<long long> d_pi_launch;   // launches of k_pi_finish that post their result to the host themselves
	long long pi_posted = 0;         // ... as counted by the host
	bool pi_direct = false;          // the last sweep enqueued posts its own result (no copy node, no stream synchronise)
	int *h_sk_dirty = nullptr;
	long long xchg_timeout_cycles = 0;
	DevBuf<double> d_erf_tab;
	int pair_grid = 0;
	DevBuf<unsigned char> d_mol_mobile;
	DevBuf<KVec> d_kvec;
	DevBuf<PairPartial> d_partials;
	DevBuf<double2> d_sk_part, d_S_mobile, d_S_frozen, d_S_all;
	DevBuf<double> d_efs, d_efi, d_efic, d_mu, d_new_mu, d_old_mu, d_rrms, d_rank, d_acc, d_dmu, d_tri, d_near, d_com, d_mol_mass, d_chain;
	DevBuf<int> d_gsctl, d_gmeta, d_nplist;
	DevBuf<double4> d_gpq;
	DevBuf<double> d_cparts, d_field_tab;
	DevBuf<int> d_mobile_sites, d_frozen_sites, d_allq, d_fp_list, d_mp_list, d_recount;
	std::vector<int> mobile_sites, frozen_sites, allq, fp_list, mp_list;
	DevBuf<unsigned long long> d_r2min_ff;
	DevBuf<double> d_t2, d_t2_cached, d_cnt_ff;
	bool rank_ff_dirty = true;
	// Gauss-Seidel pipeline: the updater kernel runs beside the solver cluster on a second stream
	cudaStream_t stream2 = nullptr;
	cudaEvent_t ev_fork = nullptr, ev_sk = nullptr, ev_pol = nullptr, ev_pre[2] = {nullptr, nullptr};
	// the Gauss-Seidel precomputation of the second sweep order (rank order) has its own buffers, so that it can be built on the second
	// stream while the first sweep runs on the first order's
	DevBuf<double4> d_gpq2;
	DevBuf<int> d_gmeta2;
	DevBuf<double> d_tri2, d_near2;
	int gs_gen = 0, gs_upd_grid = 0, gs_fused_grid = 0;   // gs_gen: generation of the pipeline's flag words (kernels_gs.cuh)
	size_t gs_ctl_len = 0;
	int *h_gs_abort = nullptr;      // pinned copy of GsCtl::abort after the last sweep of an energy()
	bool gs_ran = false;
	bool gs_fused = false;          // MPMC_GS_FUSED=1: updaters inside the solver's launch (for tools that serialise kernel launches)
	RadialTable field_tab;
	FieldParams fpar;
	std::vector<int> nplist;
	int ct_parts = 1, ct_part_len = 0;
	DevBuf<long long> d_gsprof;
	bool gs_prof_enabled = false;
	DevBuf<long long> d_pairprof;
	bool pair_prof_enabled = false;
	int pair_prof_warps = 0;
	int gs_prof_nblk = 0;
	DevBuf<unsigned long long> d_rmin;
	DevBuf<double> d_result;
	// pinned staging
	double4 *h_stage[2] = {nullptr, nullptr}; size_t stage_cap[2] = {0, 0};   // two slots: a restore and the next move may both be in flight
	bool stage_busy[2] = {false, false};
	int stage_next = 0;
	double *h_result = nullptr;
	int *h_flags = nullptr;
	// polarization bookkeeping of the last energy()
	int last_iterations = 0;
	std::vector<int> last_iters;    // per bead system
	std::vector<int> last_failed;
	bool enqueued = false;
	int gs_grid = 0;
	// bead sharding over GPUs
	nccl_comm_t comm = nullptr;
	int rank = 0, nranks = 1;
	DevBuf<double> d_pisums, d_firstcom;
	// peer-memory mailboxes of the fused all-reduce (k_pi_sums_xchg)
	bool p2p = false;
	PiMailSlot *d_mbox = nullptr;
	std::vector<PiMailSlot *> peer_mbox;
	DevBuf<PiMailSlot *> d_peers;
	DevBuf<long long> d_step;
	// the path-integral sweep (kernels + all-reduce + result copy) as a CUDA graph: one launch per Monte Carlo move
	cudaGraphExec_t pi_graph = nullptr;
	bool pi_graph_off = false, pi_warm = false;   // pi_warm: one sweep has run outside a capture (every buffer it needs exists)
	long long pi_graph_launches = 0;
	double *h_pisums = nullptr;
	// optional per-kernel-class timing with CUDA events on the engine's stream (mpmc_set_timing)
	bool timing = false;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
	std::vector<int> ev_class;          // class of every event pair used since the last collect
	size_t ev_used = 0;
	double t_ms[MPMC_NUM_KERNEL_CLASSES] = {0};
	long long t_count[MPMC_NUM_KERNEL_CLASSES] = {0};
};

namespace {
// RAII bracket: records an event pair around the kernels launched in its scope when timing is enabled
struct Timed {
	mpmc_engine *e; size_t slot = 0; bool on;
	Timed(mpmc_engine *eng, int cls) : e(eng), on(eng->timing) {
		if (!on) return;
		if (e->ev_used == e->ev_pool.size()) {
			cudaEvent_t a, b;
			cudaEventCreate(&a); cudaEventCreate(&b);
			e->ev_pool.push_back({a, b}); e->ev_class.push_back(cls);
		}
		slot = e->ev_used++;
		e->ev_class[slot] = cls;
		cudaEventRecord(e->ev_pool[slot].first, e->stream);
	}
	~Timed() { if (on) cudaEventRecord(e->ev_pool[slot].second, e->stream); }
};
void collect_timing(mpmc_engine *e) {   // call after a stream synchronize
	for (size_t i = 0; i < e->ev_used; i++) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, e->ev_pool[i].first, e->ev_pool[i].second) == cudaSuccess) {
			e->t_ms[e->ev_class[i]] += ms; e->t_count[e->ev_class[i]] += 1;
		}
	}
	e->ev_used = 0;
}
} // namespace

// ------------------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------------------
namespace {

int sync_stream(mpmc_engine *e);

// PeriodicBoundary::update (src/PeriodicBoundary.cpp:31-101) + update_pbc alphas (src/System.cpp:871-874)
int compute_cell(mpmc_engine *e, const double basis[9]) {
	CellDev &c = e->cell;
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c.b[i][j] = basis[3 * i + j];
	double (*b)[3] = c.b;
	double v = b[0][0] * (b[1][1] * b[2][2] - b[1][2] * b[2][1]);
	v += b[0][1] * (b[1][2] * b[2][0] - b[1][0] * b[2][2]);
	v += b[0][2] * (b[1][0] * b[2][1] - b[1][1] * b[2][0]);
	c.volume = v;
	if (!(v > 0)) FAIL(MPMC_ERR_INVALID_BOX, "invalid simulation box dimensions (volume %g)", v);   // System.cpp:467-470
	double short_mag = kMaxValue;
	for (int i = -15; i <= 15; i++) for (int j = -15; j <= 15; j++) for (int k = -15; k <= 15; k++) {
		if (!i && !j && !k) continue;
		double cv[3];
		for (int p = 0; p < 3; p++) cv[p] = i * b[0][p] + j * b[1][p] + k * b[2][p];
		double m = std::sqrt(cv[0] * cv[0] + cv[1] * cv[1] + cv[2] * cv[2]);
		if (m < short_mag) short_mag = m;
	}
	c.cutoff = 0.5 * short_mag;
	if (!(c.cutoff > 0)) FAIL(MPMC_ERR_INVALID_BOX, "invalid simulation box dimensions (cutoff %g)", c.cutoff);
	double iv = 1.0 / v;
	double (*r)[3] = c.rb;
	r[0][0] = iv * (b[1][1] * b[2][2] - b[1][2] * b[2][1]);
	r[0][1] = iv * (b[0][2] * b[2][1] - b[0][1] * b[2][2]);
	r[0][2] = iv * (b[0][1] * b[1][2] - b[0][2] * b[1][1]);
	r[1][0] = iv * (b[1][2] * b[2][0] - b[1][0] * b[2][2]);
	r[1][1] = iv * (b[0][0] * b[2][2] - b[0][2] * b[2][0]);
	r[1][2] = iv * (b[0][2] * b[1][0] - b[0][0] * b[1][2]);
	r[2][0] = iv * (b[1][0] * b[2][1] - b[1][1] * b[2][0]);
	r[2][1] = iv * (b[0][1] * b[2][0] - b[0][0] * b[2][1]);
	r[2][2] = iv * (b[0][0] * b[1][1] - b[0][1] * b[1][0]);
	c.ewald_alpha = e->cfg.ewald_alpha > 0 ? e->cfg.ewald_alpha : 3.5 / c.cutoff;
	c.polar_alpha = e->cfg.polar_ewald_alpha > 0 ? e->cfg.polar_ewald_alpha : 3.5 / c.cutoff;
	e->ortho = b[0][1] == 0 && b[0][2] == 0 && b[1][0] == 0 && b[1][2] == 0 && b[2][0] == 0 && b[2][1] == 0;
	// hemisphere of k vectors (src/System.Energy.cpp:1577-1590) with the two Gaussian weights precomputed
	e->kvec.clear();
	const int kmax = e->cfg.ewald_kmax;
	int l[3];
	for (l[0] = 0; l[0] <= kmax; l[0]++)
		for (l[1] = (!l[0] ? 0 : -kmax); l[1] <= kmax; l[1]++)
			for (l[2] = ((!l[0] && !l[1]) ? 1 : -kmax); l[2] <= kmax; l[2]++) {
				if (l[0] * l[0] + l[1] * l[1] + l[2] * l[2] > kmax * kmax) continue;
				KVec k;
				double kk[3];
				for (int p = 0; p < 3; p++) {
					kk[p] = 0;
					for (int q = 0; q < 3; q++) kk[p] += 2.0 * kPi * c.rb[p][q] * l[q];
				}
				k.kx = kk[0]; k.ky = kk[1]; k.kz = kk[2];
				const double k2 = kk[0] * kk[0] + kk[1] * kk[1] + kk[2] * kk[2];
				k.w_energy = std::exp(-k2 / (4.0 * c.ewald_alpha * c.ewald_alpha)) / k2;
				k.w_field = std::exp(-k2 / (4.0 * c.polar_alpha * c.polar_alpha)) / k2;
				k.l0 = l[0]; k.l1 = l[1]; k.l2 = l[2]; k.pad = 0;
				e->kvec.push_back(k);
			}
	int rc = e->d_kvec.ensure(e->kvec.size());
	if (rc) return rc;
	if (!e->kvec.empty()) CK(cudaMemcpyAsync(e->d_kvec.p, e->kvec.data(), e->kvec.size() * sizeof(KVec), cudaMemcpyHostToDevice, e->stream));
	{ int _rc = sync_stream(e); if (_rc) return _rc; }
	e->frozen_sk_dirty = true;
	e->sk_valid = false;
	e->cell_tables_dirty = true;
	e->topo_dirty = true;   // LRC and self terms depend on volume / cutoff / alpha
	return MPMC_OK;
}

// lj_lrc_corr / lj_lrc_self (src/System.Energy.cpp:1036-1096), plain-LJ branch
double lrc_formula(double eps, double sigma, double cutoff, double volume) {
	double sig_cut = std::fabs(sigma) / cutoff;
	double sig3 = std::fabs(sigma);
	sig3 *= sig3 * sig3;
	double sig_cut3 = sig_cut * sig_cut * sig_cut;
	double sig_cut9 = sig_cut3 * sig_cut3 * sig_cut3;
	return ((16.0 / 3.0) * kPi * eps * sig3) * ((1.0 / 3.0) * sig_cut9 - sig_cut3) / volume;
}

int prepare_pair_sweep(mpmc_engine *e);
int prepare_polar(mpmc_engine *e);

// Everything that depends on the site table but not on coordinates: device parameter arrays, work lists and the
// configuration-independent energy terms (pair/self LRC, Ewald point-self term).
int rebuild_topology(mpmc_engine *e) {
	const int n = e->n;
	int rc;
	if ((rc = e->d_lj.ensure(e->cap)) || (rc = e->d_alpha.ensure(e->cap)) || (rc = e->d_mass.ensure(e->cap)) || (rc = e->d_meta.ensure(e->cap))) return rc;
	std::vector<double2> lj(n);
	std::vector<int> meta(n);
	for (int i = 0; i < n; i++) {
		const bool active = e->h_eps[i] != 0.0 && e->h_sigma[i] > 0.0;   // sigma < 0 leaves pair epsilon at 0 in the reference (System.cpp:1167-1169)
		lj[i] = make_double2(active ? std::sqrt(e->h_eps[i]) : 0.0, 0.5 * e->h_sigma[i]);
		meta[i] = (e->h_mol[i] & 0x7fffffff) | (e->h_frozen[i] ? (int)0x80000000u : 0);
	}
	if (n) {
		CK(cudaMemcpyAsync(e->d_lj.p, lj.data(), n * sizeof(double2), cudaMemcpyHostToDevice, e->stream));
		CK(cudaMemcpyAsync(e->d_meta.p, meta.data(), n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
		CK(cudaMemcpyAsync(e->d_alpha.p, e->h_alpha.data(), n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
		CK(cudaMemcpyAsync(e->d_mass.p, e->h_mass.data(), n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
	}
	// work lists
	e->plist.clear(); e->mobile_q.clear(); e->frozen_q.clear(); e->mol_start.clear(); e->mol_mobile.clear();
	long long nfrozen = 0;
	for (int i = 0; i < n; i++) {
		if (e->h_alpha[i] != 0.0) e->plist.push_back(i);
		if (e->h_q[i] != 0.0) (e->h_frozen[i] ? e->frozen_q : e->mobile_q).push_back(i);
		if (i == 0 || e->h_mol[i] != e->h_mol[i - 1]) { e->mol_start.push_back(i); e->mol_mobile.push_back(e->h_frozen[i] ? 0 : 1); }
		nfrozen += e->h_frozen[i] ? 1 : 0;
	}
	e->mol_start.push_back(n);
	e->n_pair_evals = 0.5 * ((double)n * (n - 1) - (double)nfrozen * (nfrozen - 1));
	auto up = [&](auto &dbuf, const auto &vec) -> int {
		int r = dbuf.ensure(std::max<size_t>(vec.size(), 1));
		if (r) return r;
		if (!vec.empty()) CK(cudaMemcpyAsync(dbuf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice, e->stream));
		return MPMC_OK;
	};
	if ((rc = up(e->d_plist, e->plist)) || (rc = up(e->d_mobile_q, e->mobile_q)) ||
	    (rc = up(e->d_frozen_q, e->frozen_q)) || (rc = up(e->d_mol_start, e->mol_start)) ||
	    (rc = up(e->d_mol_mobile, e->mol_mobile))) return rc;
	{ int _rc = sync_stream(e); if (_rc) return _rc; }   // the std::vectors above go out of scope / may be rebuilt
	if ((rc = prepare_pair_sweep(e))) return rc;
	if ((rc = prepare_polar(e))) return rc;
	e->cell_tables_dirty = false;

	// configuration-independent terms, by (eps, sigma) type instead of by pair.  Pair LRC covers every non-frozen pair with
	// eps_ij != 0 and sigma_ij != 0, intramolecular pairs included (System.Energy.cpp:1045-1050); self LRC every non-frozen
	// site with eps, sigma != 0 (:1076-1079).
	e->lrc_pair = e->lrc_self = e->es_self = 0;
	const CellDev &c = e->cell;
	if (e->cfg.rd_lrc) {
		struct Cnt { double total = 0, frozen = 0; };
		std::map<std::pair<double, double>, Cnt> types;   // sites that can form a pair with eps_ij, sigma_ij != 0
		for (int i = 0; i < n; i++) {
			if (e->h_sigma[i] != 0 && e->h_eps[i] != 0 && !e->h_frozen[i]) e->lrc_self += lrc_formula(e->h_eps[i], e->h_sigma[i], c.cutoff, c.volume);
			if (e->h_eps[i] != 0.0 && e->h_sigma[i] > 0.0) {
				Cnt &t = types[{e->h_eps[i], e->h_sigma[i]}];
				t.total += 1; t.frozen += e->h_frozen[i] ? 1 : 0;
			}
		}
		for (auto a = types.begin(); a != types.end(); ++a)
			for (auto b = a; b != types.end(); ++b) {
				double npairs = (a == b) ? 0.5 * (a->second.total * (a->second.total - 1) - a->second.frozen * (a->second.frozen - 1))
				                         : a->second.total * b->second.total - a->second.frozen * b->second.frozen;
				if (npairs == 0) continue;
				const double eps = std::sqrt(a->first.first * b->first.first), sig = 0.5 * (a->first.second + b->first.second);
				e->lrc_pair += npairs * lrc_formula(eps, sig, c.cutoff, c.volume);
			}
	}
	if (!e->cfg.rd_only)
		for (int i = 0; i < n; i++)
			if (!e->h_frozen[i]) e->es_self -= c.ewald_alpha * e->h_q[i] * e->h_q[i] / std::sqrt(kPi);   // System.Energy.cpp:1626-1643
	e->topo_dirty = false;
	e->frozen_sk_dirty = true;
	e->sk_valid = false;
	return MPMC_OK;
}

// ---- second-generation pair sweep: host-side preparation (kernels_pair2.cuh) ----
// Largest double x in [0, hi] for which pred holds, for a predicate that is true up to some point and false beyond
// (positive doubles order like their bit patterns, so this is a bisection on the bits).
template <class P> double largest_true(P pred, double hi) {
	auto bits = [](double v) { uint64_t b; memcpy(&b, &v, 8); return b; };
	auto val = [](uint64_t b) { double v; memcpy(&v, &b, 8); return v; };
	if (pred(hi)) return hi;
	uint64_t lo = 0, up = bits(hi);
	while (up - lo > 1) {
		const uint64_t mid = lo + (up - lo) / 2;
		if (pred(val(mid))) lo = mid; else up = mid;
	}
	return val(lo);
}

int prepare_pair_sweep(mpmc_engine *e) {
	const int n = e->n;
	const CellDev &c = e->cell;
	PairParams &pp = e->pp;
	int rc2;
	const bool es = !e->cfg.rd_only;
	// What depends on the CELL only — the r^2 thresholds of the reference's cutoff tests and the erfc table (built in long double:
	// ~5 ms) — is rebuilt when the cell changes, not when the site table does: a uVT insertion or removal changes the topology on
	// every such move (and again on its rejection).
	if (e->cell_tables_dirty) {
		// the reference's cutoff tests as thresholds on r^2 (both act on rimg = sqrt(r^2), correctly rounded on the reference's host and here)
		const volatile double rc = c.cutoff;
		const double top = 4.0 * c.cutoff * c.cutoff + 1.0;
		pp.t2_lj = largest_true([&](double x) { volatile double r = std::sqrt(x); volatile double d = r - kSmallDr; return d < rc; }, top);   // System.Energy.cpp:934
		pp.t2_es = largest_true([&](double x) { volatile double r = std::sqrt(x); return !(r > rc); }, top);                                  // :1490
		pp.t2_adm = pp.t2_lj * (1.0 + 1e-9);
		pp.t2_safe = std::min(pp.t2_es, pp.t2_lj) * (1.0 - 1e-9);
		if (es) {
			const long double alpha = c.ewald_alpha;
			e->erf_tab.build(1, std::min(0.25, pp.t2_adm / 64.0), pp.t2_adm * 1.001, [&](long double u, long double *o) {
				const long double r = sqrtl(u);
				o[0] = erfcl(alpha * r) / r;
			});
			pp.u_tab_lo = e->erf_tab.u_lo; pp.tab_base = e->erf_tab.base; pp.tab_rows = e->erf_tab.nrows;
			if ((rc2 = e->d_erf_tab.ensure(e->erf_tab.rows.size()))) return rc2;
			CK(cudaMemcpyAsync(e->d_erf_tab.p, e->erf_tab.rows.data(), e->erf_tab.rows.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
		} else { pp.u_tab_lo = 0; pp.tab_base = 0; pp.tab_rows = 0; }
		pp.h_adm = (unsigned)RadialTable::hi_word(pp.t2_adm); pp.h_safe = (unsigned)RadialTable::hi_word(pp.t2_safe);
		pp.h_tab_lo = (unsigned)RadialTable::hi_word(pp.u_tab_lo);
	}
	// the class-sorted site table of the sweep: class = (frozen, charged, LJ-active); sites keep their list order inside a class
	std::vector<int> cls(n), perm(n);
	for (int i = 0; i < n; i++) {
		const bool lj_on = e->h_eps[i] != 0.0 && e->h_sigma[i] > 0.0;   // sigma < 0 leaves pair epsilon at 0 in the reference (System.cpp:1167-1169)
		cls[i] = (e->h_frozen[i] ? 4 : 0) | ((es && e->h_q[i] != 0.0) ? 2 : 0) | (lj_on ? 1 : 0);
		perm[i] = i;
	}
	std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return cls[a] < cls[b]; });
	e->perm_identity = true;
	for (int k = 0; k < n; k++) e->perm_identity = e->perm_identity && perm[k] == k;
	std::vector<int> pm(n), iperm(n);
	for (int k = 0; k < n; k++) iperm[perm[k]] = k;
	e->spq_valid = false;
	std::vector<double2> slj(n);
	int cbeg[9];
	{
		int k = 0;
		for (int cc = 0; cc < 8; cc++) { cbeg[cc] = k; while (k < n && cls[perm[k]] == cc) k++; }
		cbeg[8] = n;
	}
	for (int k = 0; k < n; k++) {
		const int i = perm[k];
		const bool lj_on = (cls[i] & 1) != 0;
		pm[k] = e->h_mol[i];
		slj[k] = make_double2(lj_on ? std::sqrt(e->h_eps[i]) : 0.0, 0.5 * e->h_sigma[i]);
	}
	// blocks of pairs: class A x class B (A <= B), what they need, none for frozen x frozen (pair->frozen, :936 / :1487)
	e->segs.clear();
	int col = 0;
	for (int ca = 0; ca < 8; ca++) {
		for (int gs = cbeg[ca]; gs < cbeg[ca + 1]; gs += 32) {
			for (int cb = ca; cb < 8; cb++) {
				if ((ca & 4) && (cb & 4)) continue;
				const int kind = ((ca & cb & 1) ? kPairLJ : 0) | ((ca & cb & 2) ? kPairES : 0);
				if (!kind) continue;
				const int jb = cb == ca ? gs + 1 : cbeg[cb], je = cbeg[cb + 1];
				if (jb >= je) continue;
				e->segs.push_back({gs, cbeg[ca + 1], jb, je, col, kind});
				col += je - jb;
			}
		}
	}
	pp.ncols = col; pp.nseg = (int)e->segs.size();
	// items: ranges of the flattened columns, dealt in ROUNDS of one item per resident warp (item index = k * B + bead, k-major).
	// The first round is assigned by warp index, the later ones through a counter.  Guided sizes: the rounds carry 1/2, 1/4, 1/8 ...
	// of the cost (the last two rounds the same), so that the sweep ends within a fraction of the SMALLEST item's duration of the
	// ideal while most of the work pays the per-item overhead (~2 k cycles) only once.  Small systems (< 128 columns per warp) get
	// a single round.  Boundaries inside a segment sit on whole 32-column chunks of that segment, so only its last chunk is partial.
	const int ctas = e->num_sms * pair_ctas_per_sm(es);
	const int warps = ctas * pair_warps(es);
	const long total_cols = (long)col * e->B;
	const double cols_per_warp = (double)total_cols / warps;
	int rounds = 1;
	while (rounds < 6 && cols_per_warp / (1 << rounds) >= 32.0) rounds++;
	if (cols_per_warp < 128.0) rounds = 1;
	if (const char *ev = getenv("MPMC_PAIR_ROUNDS")) rounds = std::max(1, atoi(ev));      // developer knob (tools/pair_time.py)
	const int per_round = std::max(1, warps / e->B);                  // items of one round in one bead system (never more items than warps)
	int K = per_round * rounds;
	if (K > std::max(1, (col + 15) / 16)) { K = std::max(1, (col + 15) / 16); rounds = 1; }    // at least ~16 columns per item
	pp.items_per_bead = K;
	e->pair_grid = std::max(1, std::min(ctas, (e->B * K + pair_warps(es) - 1) / pair_warps(es)));
	std::vector<int> item_seg(K, 0), item_col(K + 1, 0);
	{
		double total_w = 0;
		for (const PairSeg &sg : e->segs) total_w += (double)(sg.j_end - sg.j_begin) * pair_kind_weight(sg.kind);
		size_t sgi = 0;
		double before = 0;                                  // weight of the segments before sgi
		const bool align = rounds > 1;
		// cumulative share of the cost before item k
		std::vector<double> cum(K + 1, 0.0);
		for (int k = 0; k < K; k++) {
			const int r = rounds > 1 ? std::min(k / per_round, rounds - 1) : 0;
			const double share = rounds > 1 ? std::ldexp(1.0, -(std::min(r, rounds - 2) + 1)) / per_round : 1.0 / K;
			cum[k + 1] = cum[k] + share;
		}
		for (int k = 1; k < K; k++) {
			const double target = total_w * cum[k] / cum[K];
			while (sgi + 1 < e->segs.size() && before + (double)(e->segs[sgi].j_end - e->segs[sgi].j_begin) * pair_kind_weight(e->segs[sgi].kind) <= target) {
				before += (double)(e->segs[sgi].j_end - e->segs[sgi].j_begin) * pair_kind_weight(e->segs[sgi].kind);
				sgi++;
			}
			const PairSeg &sg = e->segs[sgi];
			const int len = sg.j_end - sg.j_begin;
			int within = (int)std::min<double>(len, std::max(0.0, (target - before) / pair_kind_weight(sg.kind)));
			if (align) { within = (within + 16) / 32 * 32; if (within > len) within = len; }
			item_col[k] = std::max(item_col[k - 1], sg.col0 + within);
		}
		item_col[K] = col;
		if (e->segs.empty()) std::fill(item_col.begin(), item_col.end(), 0);
	}
	for (int k = 0, sgi = 0; k < K; k++) {
		while (sgi + 1 < pp.nseg && e->segs[sgi + 1].col0 <= item_col[k]) sgi++;
		item_seg[k] = sgi;
	}
	std::vector<PairItem> items(K);
	for (int k = 0; k < K; k++) {
		items[k] = PairItem{item_col[k], item_col[k + 1], item_seg[k], 0, e->segs.empty() ? PairSeg{0, 0, 0, 0, 0, 0} : e->segs[item_seg[k]], 0, 0};
	}
	// one item per warp and nothing left for the counter: the sweep neither reads nor advances it
	pp.single_round = (e->B * K <= e->pair_grid * pair_warps(es)) ? 1 : 0;
	if ((rc2 = e->d_items.ensure(K)) || (rc2 = e->d_item_ctr.ensure(1))) return rc2;
	CK(cudaMemcpyAsync(e->d_items.p, items.data(), K * sizeof(PairItem), cudaMemcpyHostToDevice, e->stream));
	e->pair_ctr_start = e->pair_grid * pair_warps(es);
	CK(cudaMemcpyAsync(e->d_item_ctr.p, &e->pair_ctr_start, sizeof(int), cudaMemcpyHostToDevice, e->stream));
	if ((rc2 = e->d_pmeta.ensure(std::max(n, 1))) || (rc2 = e->d_segs.ensure(std::max<size_t>(e->segs.size(), 1))) || (rc2 = e->d_perm.ensure(std::max(n, 1))) || (rc2 = e->d_iperm.ensure(std::max(n, 1))) ||
	    (rc2 = e->d_slj.ensure(std::max(n, 1))) || (rc2 = e->d_spq.ensure((size_t)e->B * e->cap))) return rc2;
	CK(cudaMemcpyAsync(e->d_pmeta.p, pm.data(), n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
	CK(cudaMemcpyAsync(e->d_perm.p, perm.data(), n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
	CK(cudaMemcpyAsync(e->d_iperm.p, iperm.data(), n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
	CK(cudaMemcpyAsync(e->d_slj.p, slj.data(), n * sizeof(double2), cudaMemcpyHostToDevice, e->stream));
	if (!e->segs.empty()) CK(cudaMemcpyAsync(e->d_segs.p, e->segs.data(), e->segs.size() * sizeof(PairSeg), cudaMemcpyHostToDevice, e->stream));
	{ int _rc = sync_stream(e); if (_rc) return _rc; }
	return MPMC_OK;
}

// polarization work lists and the Thole radial table (kernels_polar2.cuh, kernels_gs.cuh)
int prepare_polar(mpmc_engine *e) {
	const mpmc_config &cf = e->cfg;
	if (!cf.polarization || cf.rd_only) return MPMC_OK;
	const int n = e->n, np = (int)e->plist.size();
	int rc;
	e->nplist.clear();
	for (int i = 0; i < n; i++) if (e->h_alpha[i] == 0.0) e->nplist.push_back(i);
	if ((rc = e->d_nplist.ensure(std::max<size_t>(e->nplist.size(), 1)))) return rc;
	if (!e->nplist.empty()) CK(cudaMemcpyAsync(e->d_nplist.p, e->nplist.data(), e->nplist.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
	// column parts of the contraction sweeps: enough (32-row block x part) CTAs for ~8 per SM, parts a multiple of the 8 column lanes
	const int row_blocks = std::max(1, (n + kOrdI - 1) / kOrdI) * e->B;
	int parts = std::max(1, std::min(32, (8 * e->num_sms + row_blocks - 1) / row_blocks));
	int len = (std::max(np, 1) + parts - 1) / parts;
	len = std::max(kOrdJ, (len + kOrdJ - 1) / kOrdJ * kOrdJ);
	parts = std::max(1, (std::max(np, 1) + len - 1) / len);
	e->ct_parts = parts; e->ct_part_len = len;
	if ((rc = e->d_cparts.ensure((size_t)32 * e->B * n * 3))) return rc;
	// static field: row / column lists and the radial table of real_term() (System.Energy.cpp:2921-2929)
	e->mobile_sites.clear(); e->frozen_sites.clear(); e->allq.clear();
	for (int i = 0; i < n; i++) {
		(e->h_frozen[i] ? e->frozen_sites : e->mobile_sites).push_back(i);
		if (e->h_q[i] != 0.0) e->allq.push_back(i);
	}
	auto upl = [&](DevBuf<int> &d, const std::vector<int> &v) -> int {
		int r2 = d.ensure(std::max<size_t>(v.size(), 1));
		if (r2) return r2;
		if (!v.empty()) CK(cudaMemcpyAsync(d.p, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
		return MPMC_OK;
	};
	e->fp_list.clear(); e->mp_list.clear();
	for (int i : e->plist) (e->h_frozen[i] ? e->fp_list : e->mp_list).push_back(i);
	if ((rc = upl(e->d_mobile_sites, e->mobile_sites)) || (rc = upl(e->d_frozen_sites, e->frozen_sites)) || (rc = upl(e->d_allq, e->allq)) ||
	    (rc = upl(e->d_fp_list, e->fp_list)) || (rc = upl(e->d_mp_list, e->mp_list))) return rc;
	if ((rc = e->d_r2min_ff.ensure(e->B)) || (rc = e->d_t2.ensure(e->B)) || (rc = e->d_t2_cached.ensure(e->B)) || (rc = e->d_recount.ensure(e->B)) ||
	    (rc = e->d_cnt_ff.ensure((size_t)e->B * n))) return rc;
	e->rank_ff_dirty = true;
	if (e->cell_tables_dirty) {
		FieldParams &fp = e->fpar;
		fp.t2_in = cf.polar_ewald ? e->pp.t2_es : e->pp.t2_lj;          // `r > rc` (:2917) / `r - 1e-12 < rc` (:3319)
		fp.t2_adm = fp.t2_in * (1.0 + 1e-9); fp.t2_safe = fp.t2_in * (1.0 - 1e-9);
		fp.u_tab_lo = 0; fp.tab_base = 0; fp.tab_rows = 0; fp.tab_len = 0;
		if (cf.polar_ewald) {
			const long double a = e->cell.polar_alpha, osp = 0.5641895835477562869480794515607725858440506293289988L;
			e->field_tab.build(2, std::min(0.25, fp.t2_adm / 64.0), fp.t2_adm * 1.001, [&](long double u, long double *o) {
				const long double r = sqrtl(u), g = 2.0L * a * osp * expl(-a * a * u) * r;
				o[0] = (g + erfcl(a * r)) / (u * r);
				o[1] = (g - erfl(a * r)) / (u * r);
			}, kTabShiftCoarse);
			fp.u_tab_lo = e->field_tab.u_lo; fp.tab_base = e->field_tab.base; fp.tab_rows = e->field_tab.nrows; fp.tab_len = (int)e->field_tab.rows.size();
			if ((rc = e->d_field_tab.ensure(e->field_tab.rows.size()))) return rc;
			CK(cudaMemcpyAsync(e->d_field_tab.p, e->field_tab.rows.data(), e->field_tab.rows.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
		}
	}
	{ int _rc = sync_stream(e); if (_rc) return _rc; }
	return MPMC_OK;
}

constexpr int kSkMaxDirty = 8;   // chunks of the mobile structure factor recomputed per evaluation before a full pass is cheaper

// every cudaStreamSynchronize of the engine's stream goes through here: whatever was staged has been consumed
int sync_stream(mpmc_engine *e) {
	CK(cudaStreamSynchronize(e->stream));
	e->stage_busy[0] = e->stage_busy[1] = false;
	return MPMC_OK;
}

int ensure_stage(mpmc_engine *e, int slot, size_t count) {
	if (count <= e->stage_cap[slot]) return MPMC_OK;
	if (e->stage_busy[0] || e->stage_busy[1]) { int rc = sync_stream(e); if (rc) return rc; }
	if (e->h_stage[slot]) cudaFreeHost(e->h_stage[slot]);
	e->h_stage[slot] = nullptr; e->stage_cap[slot] = 0;
	size_t want = std::max<size_t>(count, 1024);
	CK(cudaMallocHost(&e->h_stage[slot], want * sizeof(double4)));
	e->stage_cap[slot] = want;
	return MPMC_OK;
}

// the chunks of the mobile structure factor that hold sites [first, first + count)
void mark_moved(mpmc_engine *e, int first, int count) {
	if (!e->sk_valid) return;
	const auto lo = std::lower_bound(e->mobile_q.begin(), e->mobile_q.end(), first) - e->mobile_q.begin();
	const auto hi = std::lower_bound(e->mobile_q.begin(), e->mobile_q.end(), first + count) - e->mobile_q.begin();
	if (hi <= lo) return;
	for (int c = (int)lo / kSkSites; c <= (int)(hi - 1) / kSkSites; c++)
		if (std::find(e->sk_dirty.begin(), e->sk_dirty.end(), c) == e->sk_dirty.end()) e->sk_dirty.push_back(c);
	if ((int)e->sk_dirty.size() > kSkMaxDirty) { e->sk_valid = false; e->sk_dirty.clear(); }
}

// copy sites [first, first+count) of every bead (or one bead) from the shadow to the device: one pinned staging slot, one
// host-to-device copy, one scatter into posq (and into the class-sorted copy the pair sweep reads).  No stream synchronisation
// unless both slots are still in flight.
int push_positions(mpmc_engine *e, int bead_lo, int bead_hi, int first, int count) {
	if (count <= 0) return MPMC_OK;
	const int nb = bead_hi - bead_lo;
	const int slot = e->stage_next;
	e->stage_next ^= 1;
	int rc;
	if (e->stage_busy[slot] && (rc = sync_stream(e))) return rc;   // the slot may still be in flight from two updates ago
	if ((rc = ensure_stage(e, slot, (size_t)nb * count))) return rc;
	double4 *hs = e->h_stage[slot];
	for (int b = 0; b < nb; b++)
		for (int i = 0; i < count; i++) {
			const double *p = &e->h_pos[((size_t)(bead_lo + b) * e->n + first + i) * 3];
			hs[(size_t)b * count + i] = make_double4(p[0], p[1], p[2], e->h_q[first + i]);
		}
	if ((rc = e->d_stage.ensure((size_t)nb * count))) return rc;
	CK(cudaMemcpyAsync(e->d_stage.p, hs, (size_t)nb * count * sizeof(double4), cudaMemcpyHostToDevice, e->stream));
	e->stage_busy[slot] = true;
	const bool sorted_copy = !e->perm_identity && e->d_iperm.p && e->d_spq.p;
	k_scatter_sites<<<(nb * count + 127) / 128, 128, 0, e->stream>>>(e->d_stage.p, e->d_posq.p, e->cap, bead_lo, nb, first, count,
	                                                                 sorted_copy ? e->d_iperm.p : nullptr, sorted_copy ? e->d_spq.p : nullptr);
	e->launches++;
	CK(cudaGetLastError());
	return MPMC_OK;
}

int validate_config(const mpmc_config *cfg) {
	if (cfg->n_beads < 1) FAIL(MPMC_ERR_INVALID_SETTING, "n_beads must be >= 1");
	if (cfg->ewald_kmax < 1 || cfg->ewald_kmax > kMaxKmax) FAIL(MPMC_ERR_INVALID_SETTING, "ewald_kmax must be in [1, %d]", kMaxKmax);
	if (cfg->polarization && !cfg->rd_only) {
		// the reference's own validator for `cuda on` (src/SimulationControl.cpp:2612-2627) requires the iterative solver
		if (!cfg->polar_iterative) FAIL(MPMC_ERR_UNSUPPORTED, "GPU acceleration available for iterative Thole only, enable polar_iterative");
		if (cfg->damp_type < 0 || cfg->damp_type > 2) FAIL(MPMC_ERR_INVALID_SETTING, "Thole damping method not specified");
		if (cfg->polar_damp <= 0.0 && cfg->damp_type != MPMC_DAMPING_OFF) FAIL(MPMC_ERR_INVALID_SETTING, "damping factor must be specified");
		if (cfg->polar_precision > 0.0 && cfg->polar_max_iter > 0) FAIL(MPMC_ERR_INCOMPATIBLE, "cannot specify both polar_precision and polar_max_iter");
		if (cfg->polar_precision < 0.0) FAIL(MPMC_ERR_INVALID_SETTING, "invalid polarization iterative precision");
		if (!cfg->polar_zodid && cfg->polar_precision == 0.0 && cfg->polar_max_iter <= 0)
			FAIL(MPMC_ERR_MISSING_SETTING, "polar_max_iter or polar_precision must be set for the iterative solver");
		if (cfg->polar_sor && cfg->polar_esor) FAIL(MPMC_ERR_INCOMPATIBLE, "cannot specify both SOR and ESOR SCF methods");
	}
	return MPMC_OK;
}

template <class K> int set_smem(K kernel, size_t bytes) {
	CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
	return MPMC_OK;
}

} // namespace

static void drop_pi_graph(mpmc_engine *e) {
	if (e->pi_graph) { cudaGraphExecDestroy(e->pi_graph); e->pi_graph = nullptr; }
	e->pi_warm = false;
}

static int adopt_table(mpmc_engine *e) {
	drop_pi_graph(e);
	// (re)size device arrays for e->n sites and push everything
	const int n = e->n;
	if (n > e->cap || !e->d_posq.p) e->cap = std::max({e->cap, n + n / 2, 64});
	int rc;
	if ((rc = e->d_posq.ensure((size_t)e->B * e->cap))) return rc;
	for (int i = 1; i < n; i++)
		if (e->h_mol[i] < e->h_mol[i - 1]) FAIL(MPMC_ERR_INVALID_INPUT, "mol[] must be non-decreasing (site %d)", i);
	if ((rc = rebuild_topology(e))) return rc;
	return push_positions(e, 0, e->B, 0, n);
}


// ---- energy ----------------------------------------------------------------------------------------------
#define LAUNCHED(e) ((e)->launches++)

// chunk partials of a structure factor (all chunks, or only those listed in d_sk_dirty) and, unless `reduce` is false, their sum
static int run_structure(mpmc_engine *e, DevBuf<int> &list, int nlist, DevBuf<double2> &S, const double2 *addend, cudaStream_t stream = nullptr,
                         bool dirty_only = false, bool reduce = true) {
	if (!stream) stream = e->stream;
	const int nk = (int)e->kvec.size(), kmax = e->cfg.ewald_kmax, B = e->B;
	int rc;
	if ((rc = S.ensure((size_t)B * nk))) return rc;
	const int nchunks = (nlist + kSkSites - 1) / kSkSites;
	if (nchunks > 0) {
		if (e->d_sk_part.cap < (size_t)B * nchunks * nk) { e->sk_valid = false; dirty_only = false; }
		if ((rc = e->d_sk_part.ensure((size_t)B * nchunks * nk))) return rc;
		const size_t smem = sizeof(double2) * kSkSites * 3 * (kmax + 1);
		Timed _t(e, MPMC_K_STRUCTURE);
		if (dirty_only)
			k_structure_partial<<<dim3(std::min(kSkMaxDirty, nchunks), B, B <= 16 ? 3 : 1), kSkThreads, smem, stream>>>(e->d_posq.p, e->cap, list.p, nlist, e->d_kvec.p, nk, kmax, e->cell,
			                                                                                        e->d_sk_part.p, nchunks, e->d_sk_dirty.p);
		else
			k_structure_partial<<<dim3(nchunks, B), kSkThreads, smem, stream>>>(e->d_posq.p, e->cap, list.p, nlist, e->d_kvec.p, nk, kmax, e->cell,
			                                                                        e->d_sk_part.p, nchunks, nullptr);
		LAUNCHED(e);
	}
	if (reduce) {
		k_structure_reduce<<<dim3((nk + 127) / 128, B), 128, 0, stream>>>(e->d_sk_part.p, nchunks, nk, S.p, addend);
		LAUNCHED(e);
	}
	CK(cudaGetLastError());
	return MPMC_OK;
}

// the mobile sites' structure factor: only the chunks that hold a moved site when the stored partials are still valid (the list of
// dirty chunks travels in a pinned word array, so the same launch can sit in a CUDA graph)
static int run_structure_mobile(mpmc_engine *e, cudaStream_t stream, bool reduce) {
	const int nlist = (int)e->mobile_q.size();
	const bool inc = e->sk_valid && nlist > 0;
	if (inc) {
		e->h_sk_dirty[0] = (int)e->sk_dirty.size();
		for (size_t i = 0; i < e->sk_dirty.size(); i++) e->h_sk_dirty[1 + i] = e->sk_dirty[i];
		CK(cudaMemcpyAsync(e->d_sk_dirty.p, e->h_sk_dirty, sizeof(int) * (1 + kSkMaxDirty), cudaMemcpyHostToDevice, stream ? stream : e->stream));
	}
	int rc = run_structure(e, e->d_mobile_q, nlist, e->d_S_mobile, nullptr, stream, inc, reduce);
	if (rc) return rc;
	e->sk_valid = nlist > 0;
	e->sk_dirty.clear();
	return MPMC_OK;
}

// real-space part of the static field, added into d_efs: (mobile rows x all charged columns) + (frozen rows x mobile charged columns)
template <bool ORTHO, bool EWALD>
static int run_field_real(mpmc_engine *e) {
	const int n = e->n, B = e->B;
	const size_t smem = EWALD ? sizeof(double) * (size_t)e->fpar.tab_len : 0;
	struct Job { const int *rows; int nrows; const int *cols; int ncols; };
	const Job jobs[2] = {{e->d_mobile_sites.p, (int)e->mobile_sites.size(), e->d_allq.p, (int)e->allq.size()},
	                     {e->d_frozen_sites.p, (int)e->frozen_sites.size(), e->d_mobile_q.p, (int)e->mobile_q.size()}};
	for (const Job &j : jobs) {
		if (j.nrows == 0 || j.ncols == 0) continue;
		const int row_blocks = (j.nrows + kOrdI - 1) / kOrdI * B;
		int parts = std::max(1, std::min(32, (8 * e->num_sms + row_blocks - 1) / row_blocks));
		int len = (j.ncols + parts - 1) / parts;
		len = std::max(kOrdJ, (len + kOrdJ - 1) / kOrdJ * kOrdJ);
		parts = (j.ncols + len - 1) / len;
		k_field_parts<ORTHO, EWALD><<<dim3((j.nrows + kOrdI - 1) / kOrdI, parts, B), kOrdThreads, smem, e->stream>>>(
		    e->d_posq.p, e->d_meta.p, j.cols, j.ncols, len, j.rows, j.nrows, e->cap, e->cell, e->fpar, e->d_field_tab.p, e->d_cparts.p);
		k_field_finish<<<(j.nrows * B + 127) / 128, 128, 0, e->stream>>>(e->d_cparts.p, parts, j.rows, j.nrows, n, B, e->d_efs.p);
		e->launches += 2;
	}
	CK(cudaGetLastError());
	return MPMC_OK;
}

template <bool ORTHO>
static int run_polar(mpmc_engine *e) {
	const int n = e->n, B = e->B, np = (int)e->plist.size();
	const mpmc_config &cf = e->cfg;
	const size_t len = (size_t)B * n * 3;
	int rc;
	if ((rc = e->d_efs.ensure(len)) || (rc = e->d_efi.ensure(len)) || (rc = e->d_efic.ensure(len)) || (rc = e->d_mu.ensure(len)) ||
	    (rc = e->d_new_mu.ensure(len)) || (rc = e->d_old_mu.ensure(len)) || (rc = e->d_rrms.ensure((size_t)B * n)) ||
	    (rc = e->d_rank.ensure((size_t)B * n)) || (rc = e->d_order.ensure(std::max(np, 1))) ||
	    (rc = e->d_order.ensure(1))) return rc;
	const dim3 ogrid((n + kOrdI - 1) / kOrdI, B);
	const int nk = (int)e->kvec.size(), kmax = cf.ewald_kmax;
	// the second stream starts here, beside the static field (see the ranking block below)
	const bool side = (cf.polar_gs || cf.polar_gs_ranked) && B == 1 && !e->timing && e->stream2 && np > 0 && !cf.polar_zodid;
	if (side) {
		CK(cudaEventRecord(e->ev_pol, e->stream));
		CK(cudaStreamWaitEvent(e->stream2, e->ev_pol, 0));
	}
	// thole_field(): static field (System.Energy.cpp:3271-3296)
	if (cf.polar_ewald) {
		// S_all = S_frozen + S_mobile: the mobile chunk partials computed for coulombic_reciprocal() are still in d_sk_part
		if ((rc = e->d_S_all.ensure((size_t)B * nk))) return rc;
		k_structure_reduce<<<dim3((nk + 127) / 128, B), 128, 0, e->stream>>>(e->d_sk_part.p, ((int)e->mobile_q.size() + kSkSites - 1) / kSkSites, nk,
		                                                                     e->d_S_all.p, e->d_S_frozen.p);
		LAUNCHED(e);
		const size_t smem = sizeof(double2) * kFrSites * 3 * (kmax + 1);
{ Timed _t(e, MPMC_K_FIELD_RECIP);
		k_field_recip<<<dim3((n + kFrSites - 1) / kFrSites, B), kFrSites * kFrLanes, smem, e->stream>>>(e->d_posq.p, n, e->cap, e->d_kvec.p, nk, kmax, e->d_S_all.p,
		                                                                                     e->cell, 8.0 * kPi / e->cell.volume, e->d_efs.p);
		LAUNCHED(e);
 }		{ Timed _t(e, MPMC_K_FIELD_REAL); if ((rc = run_field_real<ORTHO, true>(e))) return rc; }
	} else {
		CK(cudaMemsetAsync(e->d_efs.p, 0, len * sizeof(double), e->stream));
		{ Timed _t(e, MPMC_K_FIELD_REAL); if ((rc = run_field_real<ORTHO, false>(e))) return rc; }
	}
	PolarDev pd;
	pd.damp = cf.polar_damp; pd.gamma = cf.polar_gamma; pd.damp_type = cf.damp_type;
	pd.gs = cf.polar_gs || cf.polar_gs_ranked; pd.sor = cf.polar_sor; pd.esor = cf.polar_esor;

	pd.allowed_sqerr = cf.polar_precision * cf.polar_precision * kDebye2Ska * kDebye2Ska;
	pd.u_damp = cf.polar_damp > 0 ? (50.0 / cf.polar_damp) * (50.0 / cf.polar_damp) : 0.0;
	const double gamma_init = (!cf.polar_sor && !cf.polar_esor) ? cf.polar_gamma : 1.0;
	const int eb = 256;
	k_dipole_init<<<(unsigned)((len + eb - 1) / eb), eb, 0, e->stream>>>(e->d_alpha.p, e->d_efs.p, n, B, gamma_init, e->d_mu.p, e->d_new_mu.p,
	                                                                    e->d_old_mu.p, e->d_efi.p, e->d_efic.p, e->d_rrms.p);
	LAUNCHED(e);
	// GS ranking (System.cpp:1000-1029) — the metric only depends on the geometry, so both sweep orders are known up front
	const bool ranked = cf.polar_gs_ranked && !cf.polar_zodid;
	// Everything the Gauss-Seidel sweeps need that depends on the geometry only — the rank metric, the two sweep orders, per order the
	// block inverses and the near tensors — is built on the SECOND stream, beside the static field, the dipole initialisation, the
	// initial contraction and (for the second order) the first sweep: k_gs_inverse runs 144 CTAs at 11 % of the FP64 pipe and the rank
	// kernels are short launches, so they fill what the main stream's kernels leave idle instead of standing in its way.
	cudaStream_t rstream = side ? e->stream2 : e->stream;
	if (ranked && np > 0) {
		Timed _t(e, MPMC_K_RANK);
		auto split = [&](int nrows, int ncols, int &parts, int &plen) {
			const int row_blocks = std::max(1, (nrows + kOrdI - 1) / kOrdI) * B;
			parts = std::max(1, std::min(32, (8 * e->num_sms + row_blocks - 1) / row_blocks));
			plen = (std::max(ncols, 1) + parts - 1) / parts;
			plen = std::max(kOrdJ, (plen + kOrdJ - 1) / kOrdJ * kOrdJ);
			parts = std::max(1, (std::max(ncols, 1) + plen - 1) / plen);
		};
		const int nfp = (int)e->fp_list.size(), nmp = (int)e->mp_list.size();
		int parts, plen;
		const unsigned long long inf_bits = 0x7ff0000000000000ull;
		if (e->rank_ff_dirty) {
			// the frozen-frozen part depends only on the frozen coordinates: once per topology / cell / framework move
			k_fill_u64<<<1, 32, 0, rstream>>>(e->d_r2min_ff.p, B, inf_bits);
			k_fill_u64<<<1, 32, 0, rstream>>>((unsigned long long *)e->d_t2_cached.p, B, 0xbff0000000000000ull);   // -1: no cached counts
			CK(cudaMemsetAsync(e->d_cnt_ff.p, 0, sizeof(double) * (size_t)B * n, rstream));
			e->launches += 2;
			if (nfp > 1) {
				split(nfp, nfp, parts, plen);
				k_rank_min_parts<ORTHO><<<dim3((nfp + kOrdI - 1) / kOrdI, parts, B), kOrdThreads, 0, rstream>>>(e->d_posq.p, e->d_fp_list.p, nfp, plen, e->d_fp_list.p, nfp,
				                                                                                                e->cap, e->cell, e->d_r2min_ff.p);
				LAUNCHED(e);
			}
			e->rank_ff_dirty = false;
		}
		k_fill_u64<<<1, 32, 0, rstream>>>(e->d_rmin.p, B, inf_bits);
		LAUNCHED(e);
		if (nmp > 0) {   // every pair with a mobile member (the frozen x mobile pairs by symmetry)
			split(nmp, np, parts, plen);
			k_rank_min_parts<ORTHO><<<dim3((nmp + kOrdI - 1) / kOrdI, parts, B), kOrdThreads, 0, rstream>>>(e->d_posq.p, e->d_plist.p, np, plen, e->d_mp_list.p, nmp,
			                                                                                                e->cap, e->cell, e->d_rmin.p);
			LAUNCHED(e);
		}
		k_rank_lim<<<(B + 31) / 32, 32, 0, rstream>>>(e->d_rmin.p, e->d_r2min_ff.p, B, e->d_t2.p, e->d_t2_cached.p, e->d_recount.p);
		k_rank_clear_gated<<<(unsigned)(((size_t)B * n + 255) / 256), 256, 0, rstream>>>(e->d_recount.p, n, B, e->d_cnt_ff.p);
		e->launches += 2;
		if (nfp > 1) {   // skipped on the device unless 1.5 rmin changed
			split(nfp, nfp, parts, plen);
			k_rank_count_parts<<<dim3((nfp + kOrdI - 1) / kOrdI, parts, B), kOrdThreads, 0, rstream>>>(e->d_posq.p, e->d_fp_list.p, nfp, plen, e->d_fp_list.p, nfp, n, e->cap,
			                                                                                           e->d_t2.p, e->d_recount.p, e->d_cnt_ff.p);
			LAUNCHED(e);
		}
		k_rank_init<<<(unsigned)(((size_t)B * n + 255) / 256), 256, 0, rstream>>>(e->d_cnt_ff.p, n, B, e->d_rank.p);
		LAUNCHED(e);
		if (nmp > 0) {
			split(nmp, np, parts, plen);
			k_rank_count_parts<<<dim3((nmp + kOrdI - 1) / kOrdI, parts, B), kOrdThreads, 0, rstream>>>(e->d_posq.p, e->d_plist.p, np, plen, e->d_mp_list.p, nmp, n, e->cap,
			                                                                                           e->d_t2.p, nullptr, e->d_rank.p);
			LAUNCHED(e);
			if (nfp > 0) {
				split(nfp, nmp, parts, plen);
				k_rank_count_parts<<<dim3((nfp + kOrdI - 1) / kOrdI, parts, B), kOrdThreads, 0, rstream>>>(e->d_posq.p, e->d_mp_list.p, nmp, plen, e->d_fp_list.p, nfp, n, e->cap,
				                                                                                           e->d_t2.p, nullptr, e->d_rank.p);
				LAUNCHED(e);
			}
		}
		CK(cudaGetLastError());
	} else CK(cudaMemsetAsync(e->d_rank.p, 0, sizeof(double) * (size_t)B * n, e->stream));

	e->last_iterations = 0;
	e->last_iters.assign(B, 0);
	std::fill(e->last_failed.begin(), e->last_failed.end(), 0);
	if (cf.polar_zodid || np == 0) return MPMC_OK;

	const bool need_old = cf.polar_rrms || cf.polar_precision > 0 || cf.polar_sor || cf.polar_esor;
	const bool want_check = cf.polar_rrms || cf.polar_precision > 0;
	const bool expd = cf.damp_type == MPMC_DAMPING_EXPONENTIAL;
	if (pd.gs && (rc = e->d_acc.ensure(len))) return rc;

	// thole_iterative() (:3450-3543) for the bead systems [b0, b0 + nb).  The reference runs it per bead system, each with its own
	// iteration count and failure flag (PathIntegral.cpp:770-780): bead systems are solved together only where that cannot change a
	// result — a fixed iteration count without Gauss-Seidel ordering.  With polar_precision, or with the Gauss-Seidel pipeline (one
	// cluster walks one system's sweep order), they are solved one after the other (nb = 1).
	auto solve = [&](int b0, int nb) -> int {
		const size_t o3 = (size_t)b0 * n * 3, o1 = (size_t)b0 * n, blen = (size_t)nb * n * 3;
		const double4 *posq = e->d_posq.p + (size_t)b0 * e->cap;
		double *mu = e->d_mu.p + o3, *new_mu = e->d_new_mu.p + o3, *old_mu = e->d_old_mu.p + o3, *efs = e->d_efs.p + o3, *efi = e->d_efi.p + o3,
		       *efic = e->d_efic.p + o3, *rrms = e->d_rrms.p + o1, *rank = e->d_rank.p + o1;
		double *acc = pd.gs ? e->d_acc.p + o3 : nullptr;
		// one contraction sweep acc_i = sum_j T_ij mu_j over a row list, then the epilogue of `mode`
		auto contract = [&](int mode, const int *rowlist, int nrows, double *out_acc) -> int {
			if (nrows <= 0) return MPMC_OK;
			const dim3 grid((nrows + kOrdI - 1) / kOrdI, e->ct_parts, nb);
			if (expd) k_contract_parts<ORTHO, true><<<grid, kOrdThreads, 0, e->stream>>>(posq, e->d_alpha.p, e->d_meta.p, e->d_plist.p, np, e->ct_part_len,
			                                                                         rowlist, nrows, n, e->cap, e->cell, pd, mu, e->d_cparts.p);
			else k_contract_parts<ORTHO, false><<<grid, kOrdThreads, 0, e->stream>>>(posq, e->d_alpha.p, e->d_meta.p, e->d_plist.p, np, e->ct_part_len,
			                                                                        rowlist, nrows, n, e->cap, e->cell, pd, mu, e->d_cparts.p);
			const int fb = 128, fg = (nrows * nb + fb - 1) / fb;
#define FINISH(M) k_contract_finish<M><<<fg, fb, 0, e->stream>>>(e->d_cparts.p, e->ct_parts, rowlist, nrows, n, nb, e->d_alpha.p, efs, efi, new_mu, efic, out_acc)
			if (mode == SWEEP_JACOBI) FINISH(SWEEP_JACOBI);
			else if (mode == SWEEP_ACC) FINISH(SWEEP_ACC);
			else FINISH(SWEEP_PALMO);
#undef FINISH
			e->launches += 2;
			CK(cudaGetLastError());
			return MPMC_OK;
		};
		int it = 0;
		bool keep = true, acc_stale = false;
		struct GsPre { DevBuf<double4> *gpq; DevBuf<int> *gmeta; DevBuf<double> *tri, *near; const int *order; };
		GsPre pre[2] = {{&e->d_gpq, &e->d_gmeta, &e->d_tri, &e->d_near, e->d_plist.p}, {&e->d_gpq2, &e->d_gmeta2, &e->d_tri2, &e->d_near2, e->d_plist.p}};
		if (!side) pre[1] = pre[0];                                  // one after the other on one stream: one set of buffers is enough
		int cur = 0;
		while (keep) {
			it++;
			if (it >= 128 && cf.polar_precision > 0) {   // MAX_ITERATION_COUNT (constants.h:52), System.Energy.cpp:3483-3494
				for (int b = b0; b < b0 + nb; b++) {
					k_dipole_fail<<<(n * 3 + eb - 1) / eb, eb, 0, e->stream>>>(e->d_alpha.p, e->d_efs.p, n, (size_t)b * n * 3, e->d_mu.p, e->d_efic.p);
					LAUNCHED(e);
					e->last_failed[b] = 1;
				}
				break;
			}
			if (need_old) CK(cudaMemcpyAsync(old_mu, mu, blen * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
			if (!pd.gs) {
				Timed _t(e, MPMC_K_DIPOLE_SWEEP);
				if ((rc = contract(SWEEP_JACOBI, e->d_plist.p, np, nullptr))) return rc;
			} else {
				// Gauss-Seidel pipeline (kernels_gs.cuh).  First sweep in list order (ranked_array = identity, :3463); from the second
				// sweep on in rank order when polar_gs_ranked (update_ranking after the first pass, :3522-3523).
				const int nblk = (np + kGsB - 1) / kGsB, nchunks = (np + kGsRows - 1) / kGsRows;
				if ((rc = e->d_dmu.ensure((size_t)np * 3)) || (rc = e->d_gsctl.ensure(sizeof(GsCtl) / sizeof(int) + nchunks))) return rc;
				if (it == 1 || acc_stale) {
					Timed _t(e, MPMC_K_DIPOLE_SWEEP);
					if ((rc = contract(SWEEP_ACC, e->d_plist.p, np, acc))) return rc;
					acc_stale = false;
				}
				// what the sweeps of one order need (k_gs_gather, k_gs_inverse, k_gs_near): order 0 = list order, order 1 = rank order
				auto precompute = [&](int k, cudaStream_t st) -> int {
					GsPre &g = pre[k];
					int r2;
					if ((r2 = g.gpq->ensure(np)) || (r2 = g.gmeta->ensure(np)) || (r2 = g.tri->ensure((size_t)nblk * kGsMat)) || (r2 = g.near->ensure((size_t)nblk * kGsNearPerBlock))) return r2;
					if (k == 1) {
						k_rank_order_plist<<<(np + kOrdI - 1) / kOrdI, kOrdThreads, 0, st>>>(rank, e->d_plist.p, np, e->d_order.p);
						LAUNCHED(e);
					}
					g.order = k == 1 ? e->d_order.p : e->d_plist.p;
					k_gs_gather<<<(np + 255) / 256, 256, 0, st>>>(posq, e->d_alpha.p, e->d_meta.p, g.order, np, g.gpq->p, g.gmeta->p);
					k_gs_inverse<ORTHO><<<nblk, kGsThreads, kGsInverseSmemBytes, st>>>(g.gpq->p, g.gmeta->p, np, e->cell, pd, g.tri->p);
					k_gs_near<ORTHO><<<nblk, kGsPipeThreads, 0, st>>>(g.gpq->p, g.gmeta->p, np, e->cell, pd, g.near->p);
					e->launches += 3;
					CK(cudaGetLastError());
					return MPMC_OK;
				};
				if (it == 1) {
					if (side) {
						// both orders on the second stream, in the order the sweeps will want them; the main stream waits per order
						if ((rc = precompute(0, e->stream2))) return rc;
						CK(cudaEventRecord(e->ev_pre[0], e->stream2));
						if (ranked) {
							if ((rc = precompute(1, e->stream2))) return rc;
							CK(cudaEventRecord(e->ev_pre[1], e->stream2));
						}
						CK(cudaStreamWaitEvent(e->stream, e->ev_pre[0], 0));
					} else {
						Timed _t(e, MPMC_K_GS_PRECOMPUTE);
						if ((rc = precompute(0, e->stream))) return rc;
					}
					cur = 0;
				} else if (ranked && it == 2) {
					if (side) CK(cudaStreamWaitEvent(e->stream, e->ev_pre[1], 0));
					else {
						Timed _t(e, MPMC_K_GS_PRECOMPUTE);
						if ((rc = precompute(1, e->stream))) return rc;
					}
					cur = 1;
				}
				const GsPre &G = pre[cur];
				const int *gs_order = G.order;
				int ns = 1;
				if (!need_old && !want_check && cf.polar_precision == 0.0) ns = (ranked && it == 1) ? 1 : (cf.polar_max_iter - it + 1);
				long long *prof = nullptr;
				if (e->gs_prof_enabled) {
					if ((rc = e->d_gsprof.ensure((size_t)nblk * 32))) return rc;
					CK(cudaMemsetAsync(e->d_gsprof.p, 0, sizeof(long long) * (size_t)nblk * 32, e->stream));
					prof = e->d_gsprof.p;
					e->gs_prof_nblk = nblk;
				}
				{
					for (int sw = 0; sw < ns; sw++) {   // one launch per sweep: the kernel boundary is the barrier between sweeps
						Timed _t(e, MPMC_K_GS_SWEEP);
						// flags count from generation << 16, so nothing is cleared between sweeps (kernels_gs.cuh); zero them when the generation wraps
						if (e->gs_gen == 0 || e->gs_gen >= (1 << (30 - kGsGenShift)) || e->gs_ctl_len != sizeof(GsCtl) / sizeof(int) + (size_t)nchunks) {
							e->gs_ctl_len = sizeof(GsCtl) / sizeof(int) + (size_t)nchunks;
							CK(cudaMemsetAsync(e->d_gsctl.p, 0, sizeof(int) * e->gs_ctl_len, e->stream));
							e->gs_gen = 0;
						}
						if (nblk >= (1 << kGsGenShift) - kGsAhead - 2) FAIL(MPMC_ERR_UNSUPPORTED, "Gauss-Seidel pipeline: too many polarizable sites (%d)", np);
						const int gbase = ++e->gs_gen << kGsGenShift;
						const int sgrid = e->gs_fused ? e->gs_fused_grid : kGsCluster;
						if (expd) k_gs_pipeline<ORTHO, true><<<sgrid, kGsPipeThreads, kGsSmemBytes, e->stream>>>(G.gpq->p, G.gmeta->p, gs_order, np, e->cell, pd, efs,
						        mu, efi, new_mu, acc, e->d_dmu.p, G.tri->p, G.near->p, (GsCtl *)e->d_gsctl.p, sw == 0 ? prof : nullptr, gbase);
						else k_gs_pipeline<ORTHO, false><<<sgrid, kGsPipeThreads, kGsSmemBytes, e->stream>>>(G.gpq->p, G.gmeta->p, gs_order, np, e->cell, pd, efs,
						        mu, efi, new_mu, acc, e->d_dmu.p, G.tri->p, G.near->p, (GsCtl *)e->d_gsctl.p, sw == 0 ? prof : nullptr, gbase);
						CK(cudaGetLastError());
						LAUNCHED(e);
						if (e->gs_fused) continue;
						// The updaters must not take the SMs the cluster needs (its CTAs want a whole SM's shared memory each): they are launched
						// as the solver kernel's programmatic dependent — same stream, eligible as soon as every CTA of the cluster has executed
						// griddepcontrol.launch_dependents, i.e. is resident.  No host wait, no second stream; the next operation in the stream
						// waits for both kernels.
						cudaLaunchConfig_t lc = {};
						lc.gridDim = dim3(e->gs_upd_grid); lc.blockDim = dim3(kGsUpdThreads); lc.dynamicSmemBytes = kGsUpdaterSmemBytes; lc.stream = e->stream;
						cudaLaunchAttribute at[1];
						at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
						lc.attrs = at; lc.numAttrs = 1;
						const double4 *a_gpq = G.gpq->p; const int *a_gmeta = G.gmeta->p; double *a_acc = acc; const double *a_dmu = e->d_dmu.p;
						GsCtl *a_ctl = (GsCtl *)e->d_gsctl.p; long long *a_prof = sw == 0 ? prof : nullptr;
						if (expd) CK(cudaLaunchKernelEx(&lc, k_gs_updaters<ORTHO, true>, a_gpq, a_gmeta, gs_order, np, e->cell, pd, a_acc, a_dmu, a_ctl, a_prof, gbase));
						else CK(cudaLaunchKernelEx(&lc, k_gs_updaters<ORTHO, false>, a_gpq, a_gmeta, gs_order, np, e->cell, pd, a_acc, a_dmu, a_ctl, a_prof, gbase));
						LAUNCHED(e);
					}
					CK(cudaGetLastError());
					CK(cudaMemcpyAsync(e->h_gs_abort, e->d_gsctl.p + 1, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
					k_gs_efi<<<(np * 3 + eb - 1) / eb, eb, 0, e->stream>>>(e->d_plist.p, np, mu, e->d_alpha.p, efs, efi);
					LAUNCHED(e);
					e->gs_ran = true;
				}
				it += ns - 1;
			}
			if (want_check) {
				CK(cudaMemsetAsync(e->d_flags.p + b0, 0, sizeof(int) * nb, e->stream));
				k_dipole_check<<<(nb * n + eb - 1) / eb, eb, 0, e->stream>>>(new_mu, old_mu, n, nb, 1, pd.allowed_sqerr, rrms, e->d_flags.p + b0);
				LAUNCHED(e);
			}
			if (cf.polar_precision == 0.0) keep = (it != cf.polar_max_iter);     // are_we_done_yet, fixed-iteration branch (:3222-3225)
			else {
				CK(cudaMemcpyAsync(e->h_flags + b0, e->d_flags.p + b0, sizeof(int) * nb, cudaMemcpyDeviceToHost, e->stream));
				{ int _rc = sync_stream(e); if (_rc) return _rc; }
				keep = false;
				for (int b = b0; b < b0 + nb; b++) keep = keep || e->h_flags[b];
			}
			if (cf.polar_palmo && !keep) {                                       // :3518-3519
				Timed _t(e, MPMC_K_PALMO);
				if (pd.gs) {   // the running contraction already is sum_j T_ij mu_j for the polarizable rows
					k_gs_palmo<<<(np * 3 + eb - 1) / eb, eb, 0, e->stream>>>(e->d_plist.p, np, efi, acc, efic);
					LAUNCHED(e);
					if ((rc = contract(SWEEP_PALMO_NONPOLAR, e->d_nplist.p, (int)e->nplist.size(), nullptr))) return rc;
				} else {
					if ((rc = contract(SWEEP_PALMO, nullptr, n, nullptr))) return rc;
				}
			}
			if (!pd.gs || cf.polar_sor || cf.polar_esor) {                       // :3526-3536 (plain GS already has mu == new_mu)
				k_mu_update<<<(unsigned)((blen + eb - 1) / eb), eb, 0, e->stream>>>(new_mu, old_mu, blen, cf.polar_sor, cf.polar_esor, cf.polar_gamma,
				                                                                   std::exp(-cf.polar_gamma * it), mu);
				LAUNCHED(e);
				if (pd.gs) acc_stale = true;                                     // relaxation moved mu: the running contraction must be rebuilt
			}
		}
		for (int b = b0; b < b0 + nb; b++) e->last_iters[b] = it;
		return MPMC_OK;
	};
	if (B > 1 && (pd.gs || cf.polar_precision > 0)) {
		for (int b = 0; b < B; b++) if ((rc = solve(b, 1))) return rc;
	} else if ((rc = solve(0, B))) return rc;
	e->last_iterations = e->last_iters[0];
	CK(cudaGetLastError());
	return MPMC_OK;
}

// pi_fused: the path-integral aggregate is wanted, not the per-bead records — the reductions, the reciprocal energy, the per-bead
// assembly and the cross-GPU exchange run as one kernel (k_pi_finish) and the per-bead result copy is skipped
template <bool ORTHO>
static int enqueue_energy(mpmc_engine *e, bool pi_fused = false) {
	const int n = e->n, B = e->B;
	const mpmc_config &cf = e->cfg;
	int rc;
	if (n < 1) FAIL(MPMC_ERR_NO_MOLECULES, "energy: no sites uploaded");
	if (e->topo_dirty && (rc = rebuild_topology(e))) return rc;
	Timed _whole(e, MPMC_K_ENERGY_TOTAL);
	const bool es = !cf.rd_only;
	// coulombic_reciprocal(), structure factors.  The framework's: only when the cell or a frozen charged site changed.  The mobile
	// sites': on the second stream, beside the pair sweep — it needs the coordinates only, and the sweep (whose items are dealt through
	// a counter) takes whatever the structure-factor CTAs leave free, so its ramp-up and tail are no longer idle time.
	const bool sk_side = es && !e->timing && e->stream2 && e->ev_fork && e->ev_sk;
	pi_fused = pi_fused && !cf.polarization;
	if (es && e->frozen_sk_dirty) {
		if ((rc = run_structure(e, e->d_frozen_q, (int)e->frozen_q.size(), e->d_S_frozen, nullptr))) return rc;
		e->frozen_sk_dirty = false;
		e->sk_valid = false;                 // the framework pass used the same scratch for its chunk partials
	}
	if (sk_side) {
		CK(cudaEventRecord(e->ev_fork, e->stream));
		CK(cudaStreamWaitEvent(e->stream2, e->ev_fork, 0));
		if ((rc = run_structure_mobile(e, e->stream2, !pi_fused))) return rc;
		CK(cudaEventRecord(e->ev_sk, e->stream2));
	}
	// pair sweep: lj() + coulombic_real()
	const int nitems = e->pp.items_per_bead;
	{
		if ((rc = e->d_partials.ensure((size_t)B * std::max(nitems, 1)))) return rc;
		const size_t smem = pair_sweep_smem(es, e->pp.tab_rows);
		Timed _t(e, MPMC_K_PAIR);
		const double4 *spq = e->d_posq.p;        // one class in list order (bulk LJ, single-site models): the table already is sorted
		if (!e->perm_identity) {
			if (!e->spq_valid) {             // after a topology change; moves keep the copy current (k_scatter_sites)
				k_pair_gather<<<(unsigned)(((size_t)B * n + 255) / 256), 256, 0, e->stream>>>(e->d_posq.p, e->d_perm.p, n, e->cap, B, e->d_spq.p);
				LAUNCHED(e);
				e->spq_valid = true;
			}
			spq = e->d_spq.p;
		}
		long long *pprof = nullptr;
		if (e->pair_prof_enabled) {
			e->pair_prof_warps = e->pair_grid * pair_warps(es);
			if ((rc = e->d_pairprof.ensure((size_t)4 * e->pair_prof_warps))) return rc;
			CK(cudaMemsetAsync(e->d_pairprof.p, 0, sizeof(long long) * 4 * (size_t)e->pair_prof_warps, e->stream));
			pprof = e->d_pairprof.p;
		}
		if (es) k_pair_sweep<ORTHO, true><<<e->pair_grid, pair_warps(true) * 32, smem, e->stream>>>(spq, e->d_slj.p, e->d_pmeta.p, e->cap, B, e->d_segs.p, e->d_items.p, e->pp, e->cell, e->d_erf_tab.p, e->d_partials.p, e->d_item_ctr.p, pprof);
		else k_pair_sweep<ORTHO, false><<<e->pair_grid, pair_warps(false) * 32, smem, e->stream>>>(spq, e->d_slj.p, e->d_pmeta.p, e->cap, B, e->d_segs.p, e->d_items.p, e->pp, e->cell, nullptr, e->d_partials.p, e->d_item_ctr.p, pprof);
		LAUNCHED(e);
	}
	if (pi_fused) {
		const int nk = (int)e->kvec.size();
		if (es) {
			if (sk_side) CK(cudaStreamWaitEvent(e->stream, e->ev_sk, 0));
			else if ((rc = run_structure_mobile(e, nullptr, false))) return rc;
		}
		if ((rc = e->d_pisums.ensure(8 + 4 * (size_t)B)) || (rc = e->d_pi_done.ensure(1))) return rc;
		const bool xchg = e->p2p;
		// single GPU or peer-memory exchange: the kernel's sums are final and it posts them to the host itself; with the NCCL
		// fallback the all-reduce still follows and the result is copied afterwards
		e->pi_direct = xchg || !e->comm;
		k_pi_finish<<<B, 256, 0, e->stream>>>(e->d_partials.p, nitems, e->d_item_ctr.p, e->pair_ctr_start, e->d_sk_part.p,
		                                      ((int)e->mobile_q.size() + kSkSites - 1) / kSkSites, e->d_kvec.p, nk, 4.0 * kPi / e->cell.volume, e->d_S_mobile.p,
		                                      e->d_result.p, B, e->lrc_pair + e->lrc_self, e->es_self, es ? 1 : 0, e->d_pisums.p,
		                                      xchg ? e->d_peers.p : nullptr, e->rank, e->nranks, xchg ? e->d_step.p : nullptr, e->xchg_timeout_cycles, e->d_pi_done.p,
		                                      e->pi_direct ? e->h_pisums : nullptr, e->d_pi_launch.p);
		LAUNCHED(e);
		CK(cudaGetLastError());
		e->enqueued = true;
		return MPMC_OK;
	}
	k_reduce_partials<<<B, 256, 0, e->stream>>>(e->d_partials.p, nitems, e->d_result.p, e->d_item_ctr.p, e->pair_ctr_start);
	LAUNCHED(e);
	if (es) {
		const int nk = (int)e->kvec.size();
		if (sk_side) CK(cudaStreamWaitEvent(e->stream, e->ev_sk, 0));
		else if ((rc = run_structure_mobile(e, nullptr, true))) return rc;
		k_recip_energy<<<B, 256, 0, e->stream>>>(e->d_S_mobile.p, e->d_kvec.p, nk, 4.0 * kPi / e->cell.volume, e->d_result.p);
		LAUNCHED(e);
		if (cf.polarization) {
			if ((rc = run_polar<ORTHO>(e))) return rc;
			k_polar_energy<<<B, 1024, 0, e->stream>>>(e->d_mu.p, e->d_efs.p, e->d_efic.p, e->d_rrms.p, n, e->d_result.p);
			LAUNCHED(e);
		}
	}
	CK(cudaMemcpyAsync(e->h_result, e->d_result.p, sizeof(double) * res_len(B), cudaMemcpyDeviceToHost, e->stream));
	CK(cudaGetLastError());
	e->enqueued = true;
	return MPMC_OK;
}


// ------------------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" {

int mpmc_abi_version(void) { return MPMC_ABI_VERSION; }
const char *mpmc_last_error(void) { return g_err.c_str(); }

int mpmc_device_count(int *count) {
	*count = 0;
	CK(cudaGetDeviceCount(count));
	return MPMC_OK;
}

int mpmc_create(const mpmc_config *cfg, mpmc_engine **out) {
	*out = nullptr;
	int rc = validate_config(cfg);
	if (rc) return rc;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		FAIL(MPMC_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
	}
	if (cfg->device < 0 || cfg->device >= ndev) FAIL(MPMC_ERR_INVALID_SETTING, "device %d out of range (%d devices)", cfg->device, ndev);
	CK(cudaSetDevice(cfg->device));
	mpmc_engine *e = new mpmc_engine();
	e->cfg = *cfg;
	e->B = cfg->n_beads;
	e->dev = cfg->device;
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, cfg->device));
	e->num_sms = prop.multiProcessorCount;
	CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
	CK(cudaMallocHost(&e->h_result, sizeof(double) * res_len(e->B)));
	CK(cudaMallocHost(&e->h_flags, sizeof(int) * e->B));
	if ((rc = e->d_result.ensure(res_len(e->B))) || (rc = e->d_flags.ensure(e->B)) || (rc = e->d_rmin.ensure(e->B))) { mpmc_destroy(e); return rc; }
	e->last_failed.assign(e->B, 0);
	CK(cudaMallocHost(&e->h_sk_dirty, sizeof(int) * (1 + kSkMaxDirty)));
	if ((rc = e->d_sk_dirty.ensure(1 + kSkMaxDirty)) || (rc = e->d_pi_done.ensure(1)) || (rc = e->d_pisums.ensure(8 + 4 * (size_t)e->B)) ||
	    (rc = e->d_pi_launch.ensure(1))) { mpmc_destroy(e); return rc; }
	CK(cudaMemset(e->d_pi_done.p, 0, sizeof(int)));
	CK(cudaMemset(e->d_pi_launch.p, 0, sizeof(long long)));
	CK(cudaMallocHost(&e->h_pisums, sizeof(double) * (8 + 4 * (size_t)e->B)));
	memset(e->h_pisums, 0, sizeof(double) * (8 + 4 * (size_t)e->B));
	CK(cudaMemset(e->d_pisums.p, 0, sizeof(double) * 8));
	{
		const char *ts = getenv("MPMC_PI_XCHG_TIMEOUT_S");
		const double sec = ts ? atof(ts) : 60.0;
		e->xchg_timeout_cycles = (long long)(std::max(sec, 0.001) * 1.0e3 * prop.clockRate);     // clockRate is in kHz
	}
	// shared-memory opt-ins
	if ((rc = set_smem(k_structure_partial, sizeof(double2) * kSkSites * 3 * (kMaxKmax + 1))) ||
	    (rc = set_smem(k_field_recip, sizeof(double2) * kFrSites * 3 * (kMaxKmax + 1))) ||
	    false) { mpmc_destroy(e); return rc; }
	{
		const size_t fmax = 96 * 1024;
		if ((rc = set_smem(k_field_parts<true, true>, fmax)) || (rc = set_smem(k_field_parts<false, true>, fmax))) { mpmc_destroy(e); return rc; }
		if ((rc = set_smem(k_gs_inverse<true>, kGsInverseSmemBytes)) || (rc = set_smem(k_gs_inverse<false>, kGsInverseSmemBytes))) { mpmc_destroy(e); return rc; }
		if ((rc = set_smem(k_gs_pipeline<true, true>, kGsSmemBytes)) || (rc = set_smem(k_gs_pipeline<false, true>, kGsSmemBytes)) ||
		    (rc = set_smem(k_gs_pipeline<true, false>, kGsSmemBytes)) || (rc = set_smem(k_gs_pipeline<false, false>, kGsSmemBytes))) { mpmc_destroy(e); return rc; }
	}
	{
		const size_t pmax = pair_sweep_smem(true, 1024);     // tables of up to 1024 rows (32 octaves)
		if ((rc = set_smem(k_pair_sweep<true, true>, pmax)) || (rc = set_smem(k_pair_sweep<false, true>, pmax)) ||
		    (rc = set_smem(k_pair_sweep<true, false>, pmax)) || (rc = set_smem(k_pair_sweep<false, false>, pmax))) { mpmc_destroy(e); return rc; }
	}
	{
		// the Gauss-Seidel pipeline: one cluster (solver + helpers) and an updater kernel on every other SM, 2 CTAs each
		CK(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
		CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&e->ev_sk, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&e->ev_pol, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&e->ev_pre[0], cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&e->ev_pre[1], cudaEventDisableTiming));
		CK(cudaMallocHost(&e->h_gs_abort, sizeof(int)));
		*e->h_gs_abort = 0;
		if ((rc = set_smem(k_gs_updaters<true, true>, kGsUpdaterSmemBytes)) || (rc = set_smem(k_gs_updaters<false, true>, kGsUpdaterSmemBytes)) ||
		    (rc = set_smem(k_gs_updaters<true, false>, kGsUpdaterSmemBytes)) || (rc = set_smem(k_gs_updaters<false, false>, kGsUpdaterSmemBytes))) { mpmc_destroy(e); return rc; }
		int occ = 0;
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gs_updaters<true, true>, kGsUpdThreads, kGsUpdaterSmemBytes));
		if (occ < 1) FAIL(MPMC_ERR_CUDA, "the Gauss-Seidel updater kernel does not fit an SM");
		e->gs_upd_grid = std::max(1, e->num_sms - kGsCluster);
		e->gs_grid = kGsCluster;
		// single-launch fallback: as many clusters as the device holds at one CTA per SM
		cudaLaunchConfig_t lc = {};
		lc.gridDim = dim3(e->num_sms / kGsCluster * kGsCluster); lc.blockDim = dim3(kGsPipeThreads); lc.dynamicSmemBytes = kGsSmemBytes;
		cudaLaunchAttribute at[1];
		at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = kGsCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
		lc.attrs = at; lc.numAttrs = 1;
		int ncl = 0;
		CK(cudaOccupancyMaxActiveClusters(&ncl, k_gs_pipeline<true, true>, &lc));
		e->gs_fused_grid = std::max(2, ncl) * kGsCluster;
		const char *fz = getenv("MPMC_GS_FUSED");
		e->gs_fused = fz && fz[0] == '1';
	}
	if ((rc = compute_cell(e, cfg->basis))) { mpmc_destroy(e); return rc; }
	if (cfg->capacity > 0) e->cap = cfg->capacity;
	*out = e;
	return MPMC_OK;
}

int mpmc_destroy(mpmc_engine *e) {
	if (!e) return MPMC_OK;
	cudaSetDevice(e->dev);
	if (e->stream) cudaStreamSynchronize(e->stream);
	e->d_posq.release(); e->d_lj.release(); e->d_alpha.release(); e->d_mass.release(); e->d_meta.release(); e->d_plist.release();
	e->d_mobile_q.release(); e->d_frozen_q.release(); e->d_mol_start.release(); e->d_order.release(); e->d_flags.release();
	e->d_segs.release(); e->d_items.release(); e->d_item_ctr.release(); e->d_pairprof.release(); e->d_pmeta.release(); e->d_perm.release(); e->d_slj.release(); e->d_spq.release(); e->d_stage.release(); e->d_erf_tab.release(); e->d_mol_mobile.release(); e->d_kvec.release(); e->d_partials.release();
	e->d_sk_part.release(); e->d_S_mobile.release(); e->d_S_frozen.release(); e->d_S_all.release();
	e->d_efs.release(); e->d_efi.release(); e->d_efic.release(); e->d_mu.release(); e->d_new_mu.release(); e->d_old_mu.release();
	e->d_rrms.release(); e->d_rank.release(); e->d_acc.release(); e->d_dmu.release(); e->d_tri.release(); e->d_near.release(); e->d_gsctl.release(); e->d_gmeta.release(); e->d_nplist.release(); e->d_gpq.release(); e->d_cparts.release(); e->d_field_tab.release(); e->d_fp_list.release(); e->d_mp_list.release(); e->d_recount.release(); e->d_r2min_ff.release(); e->d_t2.release(); e->d_t2_cached.release(); e->d_cnt_ff.release(); e->d_mobile_sites.release(); e->d_frozen_sites.release(); e->d_allq.release(); e->d_com.release(); e->d_mol_mass.release(); e->d_chain.release();
	e->d_rmin.release(); e->d_result.release();
	for (int q = 0; q < 2; q++) if (e->h_stage[q]) cudaFreeHost(e->h_stage[q]);
	if (e->h_sk_dirty) cudaFreeHost(e->h_sk_dirty);
	e->d_iperm.release(); e->d_sk_dirty.release(); e->d_pi_done.release(); e->d_pi_launch.release();
	if (e->h_result) cudaFreeHost(e->h_result);
	if (e->h_flags) cudaFreeHost(e->h_flags);
	drop_pi_graph(e);
	if (e->stream2) { cudaStreamSynchronize(e->stream2); cudaStreamDestroy(e->stream2); }
	if (e->ev_fork) cudaEventDestroy(e->ev_fork);
	if (e->ev_sk) cudaEventDestroy(e->ev_sk);
	if (e->ev_pol) cudaEventDestroy(e->ev_pol);
	for (int q = 0; q < 2; q++) if (e->ev_pre[q]) cudaEventDestroy(e->ev_pre[q]);
	e->d_gpq2.release(); e->d_gmeta2.release(); e->d_tri2.release(); e->d_near2.release();
	if (e->h_gs_abort) cudaFreeHost(e->h_gs_abort);
	for (int r = 0; r < (int)e->peer_mbox.size(); r++) if (r != e->rank && e->peer_mbox[r]) cudaIpcCloseMemHandle(e->peer_mbox[r]);
	if (e->d_mbox) cudaFree(e->d_mbox);
	e->d_peers.release(); e->d_step.release();
	if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
	e->d_pisums.release(); e->d_firstcom.release();
	if (e->h_pisums) cudaFreeHost(e->h_pisums);
	for (auto &p : e->ev_pool) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
	if (e->stream) cudaStreamDestroy(e->stream);
	delete e;
	return MPMC_OK;
}

int mpmc_set_cell(mpmc_engine *e, const double basis[9]) {
	CK(cudaSetDevice(e->dev));
	drop_pi_graph(e);
	memcpy(e->cfg.basis, basis, sizeof(double) * 9);
	return compute_cell(e, basis);
}

int mpmc_get_cell(mpmc_engine *e, double out[22]) {
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { out[3 * i + j] = e->cell.b[i][j]; out[9 + 3 * i + j] = e->cell.rb[i][j]; }
	out[18] = e->cell.volume; out[19] = e->cell.cutoff; out[20] = e->cell.ewald_alpha; out[21] = e->cell.polar_alpha;
	return MPMC_OK;
}

int mpmc_num_sites(mpmc_engine *e, int *n) { *n = e->n; return MPMC_OK; }

int mpmc_upload_sites(mpmc_engine *e, int n, const double *pos, const double *charge, const double *alpha,
                      const double *epsilon, const double *sigma, const double *mass, const int *mol, const int *frozen) {
	CK(cudaSetDevice(e->dev));
	if (n < 1) FAIL(MPMC_ERR_NO_MOLECULES, "no sites");
	bool mobile = false;
	for (int i = 0; i < n; i++) mobile = mobile || !frozen[i];
	if (!mobile) FAIL(6000 /* missing_required_datum, System.cpp:757-760 */, "no moveable molecules found");
	e->n = n;
	e->h_pos.assign(pos, pos + (size_t)e->B * n * 3);
	e->h_q.assign(charge, charge + n); e->h_alpha.assign(alpha, alpha + n); e->h_eps.assign(epsilon, epsilon + n);
	e->h_sigma.assign(sigma, sigma + n); e->h_mass.assign(mass, mass + n);
	e->h_mol.assign(mol, mol + n); e->h_frozen.assign(frozen, frozen + n);
	return adopt_table(e);
}

int mpmc_update_sites(mpmc_engine *e, int bead, int first, int count, const double *pos) {
	CK(cudaSetDevice(e->dev));
	if (bead < 0 || bead >= e->B || first < 0 || count < 0 || first + count > e->n) FAIL(MPMC_ERR_INVALID_INPUT, "update_sites: range out of bounds");
	memcpy(&e->h_pos[((size_t)bead * e->n + first) * 3], pos, sizeof(double) * 3 * count);
	for (int i = first; i < first + count; i++) if (e->h_frozen[i]) { e->rank_ff_dirty = true; if (e->h_q[i] != 0.0) e->frozen_sk_dirty = true; }
	mark_moved(e, first, count);
	return push_positions(e, bead, bead + 1, first, count);
}

int mpmc_update_sites_all_beads(mpmc_engine *e, int first, int count, const double *pos) {
	CK(cudaSetDevice(e->dev));
	if (first < 0 || count < 0 || first + count > e->n) FAIL(MPMC_ERR_INVALID_INPUT, "update_sites: range out of bounds");
	for (int b = 0; b < e->B; b++) memcpy(&e->h_pos[((size_t)b * e->n + first) * 3], pos + (size_t)b * count * 3, sizeof(double) * 3 * count);
	for (int i = first; i < first + count; i++) if (e->h_frozen[i]) { e->rank_ff_dirty = true; if (e->h_q[i] != 0.0) e->frozen_sk_dirty = true; }
	mark_moved(e, first, count);
	return push_positions(e, 0, e->B, first, count);
}

int mpmc_insert_sites(mpmc_engine *e, int before, int count, const double *pos, const double *charge, const double *alpha,
                      const double *epsilon, const double *sigma, const double *mass, int frozen) {
	CK(cudaSetDevice(e->dev));
	if (before < 0 || before > e->n || count < 1) FAIL(MPMC_ERR_INVALID_INPUT, "insert_sites: bad position");
	const int n0 = e->n, n1 = n0 + count;
	// the inserted sites form one new molecule; molecule indices after it shift up by one
	const int newmol = before < n0 ? e->h_mol[before] : (n0 ? e->h_mol[n0 - 1] + 1 : 0);
	if (before > 0 && before < n0 && e->h_mol[before - 1] == e->h_mol[before]) FAIL(MPMC_ERR_INVALID_INPUT, "insert_sites: position splits a molecule");
	std::vector<double> np((size_t)e->B * n1 * 3);
	for (int b = 0; b < e->B; b++) {
		const double *src = &e->h_pos[(size_t)b * n0 * 3];
		double *dst = &np[(size_t)b * n1 * 3];
		memcpy(dst, src, sizeof(double) * 3 * before);
		memcpy(dst + 3 * before, pos + (size_t)b * count * 3, sizeof(double) * 3 * count);
		memcpy(dst + 3 * (before + count), src + 3 * before, sizeof(double) * 3 * (n0 - before));
	}
	e->h_pos.swap(np);
	auto ins = [&](std::vector<double> &v, const double *src) { v.insert(v.begin() + before, src, src + count); };
	ins(e->h_q, charge); ins(e->h_alpha, alpha); ins(e->h_eps, epsilon); ins(e->h_sigma, sigma); ins(e->h_mass, mass);
	for (int i = before; i < n0; i++) e->h_mol[i] += 1;
	e->h_mol.insert(e->h_mol.begin() + before, count, newmol);
	e->h_frozen.insert(e->h_frozen.begin() + before, count, frozen ? 1 : 0);
	e->n = n1;
	return adopt_table(e);
}

int mpmc_remove_sites(mpmc_engine *e, int first, int count) {
	CK(cudaSetDevice(e->dev));
	if (first < 0 || count < 1 || first + count > e->n) FAIL(MPMC_ERR_INVALID_INPUT, "remove_sites: range out of bounds");
	if (count == e->n) FAIL(MPMC_ERR_NO_MOLECULES, "remove_sites: would empty the system");
	const int n0 = e->n, n1 = n0 - count;
	std::vector<double> np((size_t)e->B * n1 * 3);
	for (int b = 0; b < e->B; b++) {
		const double *src = &e->h_pos[(size_t)b * n0 * 3];
		double *dst = &np[(size_t)b * n1 * 3];
		memcpy(dst, src, sizeof(double) * 3 * first);
		memcpy(dst + 3 * first, src + 3 * (first + count), sizeof(double) * 3 * (n0 - first - count));
	}
	e->h_pos.swap(np);
	auto del = [&](auto &v) { v.erase(v.begin() + first, v.begin() + first + count); };
	del(e->h_q); del(e->h_alpha); del(e->h_eps); del(e->h_sigma); del(e->h_mass); del(e->h_mol); del(e->h_frozen);
	e->n = n1;
	return adopt_table(e);
}

int mpmc_energy_enqueue(mpmc_engine *e) {
	CK(cudaSetDevice(e->dev));
	return e->ortho ? enqueue_energy<true>(e) : enqueue_energy<false>(e);
}

int mpmc_energy_fetch(mpmc_engine *e, mpmc_energy_out *out) {
	if (!e->enqueued) FAIL(MPMC_ERR_INTERNAL, "energy_fetch without energy_enqueue");
	CK(cudaSetDevice(e->dev));
	{ int _rc = sync_stream(e); if (_rc) return _rc; }
	e->enqueued = false;
	if (e->timing) collect_timing(e);
	if (e->gs_ran) {
		e->gs_ran = false;
		if (*e->h_gs_abort) {
			// the solver and updater kernels were not run side by side (a tool that serialises kernel launches): the result of this
			// evaluation is meaningless.  Switch this engine to the single-launch pipeline for good and evaluate again.
			*e->h_gs_abort = 0;
			e->gs_gen = 0;                                  // the sticky abort word is zeroed with the other flags before the next sweep
			if (e->gs_fused) FAIL(MPMC_ERR_CUDA, "the Gauss-Seidel pipeline timed out waiting for its own CTAs");
			e->gs_fused = true;
			int rc = mpmc_energy_enqueue(e);
			if (rc) return rc;
			return mpmc_energy_fetch(e, out);
		}
	}
	const int B = e->B;
	const mpmc_config &cf = e->cfg;
	for (int b = 0; b < B; b++) {
		mpmc_energy_out &o = out[b];
		memset(&o, 0, sizeof o);
		const double *res = e->h_result + (size_t)kResStride * b;
		const double rd = res[0], re = res[1], in = res[2], cnt = res[3];
		o.rd_pair = rd; o.rd_lrc_pair = e->lrc_pair; o.rd_lrc_self = e->lrc_self;
		o.rd_energy = rd + e->lrc_pair + e->lrc_self;
		o.n_pairs_in_cutoff = cnt; o.n_pair_evals = e->n_pair_evals;
		if (!cf.rd_only) {
			o.es_real = re; o.es_self_intra = in; o.es_reciprocal = res[4]; o.es_self = e->es_self;
			o.coulombic_energy = (re - in) + o.es_reciprocal + o.es_self;     // System.Energy.cpp:1407-1412, :1510
			if (cf.polarization) {
				const double *pr = res + 5;
				double pot = pr[0];
				if (cf.polar_palmo) pot += pr[1];
				o.polarization_energy = -0.5 * pot;                           // :2609-2618
				o.dipole_rrms = pr[2] / e->n;               // :2639-2657
				o.polarization_iterations = b < (int)e->last_iters.size() ? e->last_iters[b] : e->last_iterations;
				o.iterator_failed = e->last_failed[b];
			}
		}
		o.energy = o.rd_energy + o.coulombic_energy + o.polarization_energy + o.vdw_energy;   // :136
	}
	return MPMC_OK;
}

int mpmc_energy(mpmc_engine *e, mpmc_energy_out *out) {
	int rc = mpmc_energy_enqueue(e);
	if (rc) return rc;
	return mpmc_energy_fetch(e, out);
}

int mpmc_download_dipoles(mpmc_engine *e, int bead, double *mu, double *ef_static, double *ef_induced, double *ef_induced_change) {
	CK(cudaSetDevice(e->dev));
	if (bead < 0 || bead >= e->B) FAIL(MPMC_ERR_INVALID_INPUT, "bead out of range");
	if (!e->d_mu.p) FAIL(MPMC_ERR_INVALID_SETTING, "polarization has not been evaluated");
	const size_t len = (size_t)e->n * 3, off = (size_t)bead * len;
	{ int _rc = sync_stream(e); if (_rc) return _rc; }
	if (mu) CK(cudaMemcpy(mu, e->d_mu.p + off, len * sizeof(double), cudaMemcpyDeviceToHost));
	if (ef_static) CK(cudaMemcpy(ef_static, e->d_efs.p + off, len * sizeof(double), cudaMemcpyDeviceToHost));
	if (ef_induced) CK(cudaMemcpy(ef_induced, e->d_efi.p + off, len * sizeof(double), cudaMemcpyDeviceToHost));
	if (ef_induced_change) CK(cudaMemcpy(ef_induced_change, e->d_efic.p + off, len * sizeof(double), cudaMemcpyDeviceToHost));
	return MPMC_OK;
}

int mpmc_download_rank_metric(mpmc_engine *e, int bead, double *rank_metric) {
	CK(cudaSetDevice(e->dev));
	if (bead < 0 || bead >= e->B) FAIL(MPMC_ERR_INVALID_INPUT, "bead out of range");
	if (!e->d_rank.p) FAIL(MPMC_ERR_INVALID_SETTING, "polarization has not been evaluated");
	{ int _rc = sync_stream(e); if (_rc) return _rc; }
	CK(cudaMemcpy(rank_metric, e->d_rank.p + (size_t)bead * e->n, e->n * sizeof(double), cudaMemcpyDeviceToHost));
	return MPMC_OK;
}

// enqueue one sweep over the local bead systems and leave {sum rd, sum coulombic, sum polarization, sum vdw} in d_pisums[0..3]
// `fused` (the all-reduce path): one kernel does the reductions, the assembly and the exchange (k_pi_finish); polarizable systems
// and the per-bead API keep the separate kernels.
static int pi_sums_enqueue(mpmc_engine *e, double *d_per_bead, bool xchg = false, bool fused = false) {
	if (!e->h_pisums) CK(cudaMallocHost(&e->h_pisums, sizeof(double) * (8 + 4 * (size_t)e->B)));
	const mpmc_config &cf = e->cfg;
	if (fused && !cf.polarization) {
		CK(cudaSetDevice(e->dev));
		return e->ortho ? enqueue_energy<true>(e, true) : enqueue_energy<false>(e, true);
	}
	int rc = mpmc_energy_enqueue(e);
	if (rc) return rc;
	if (e->p2p && xchg)
		k_pi_sums_xchg<<<1, 32, 0, e->stream>>>(e->d_result.p, e->B, e->lrc_pair + e->lrc_self, e->es_self, !cf.rd_only, cf.polarization, cf.polar_palmo,
		                                        e->d_pisums.p, e->d_peers.p, e->rank, e->nranks, e->d_step.p, e->xchg_timeout_cycles);
	else
		k_pi_sums<<<1, 32, 0, e->stream>>>(e->d_result.p, e->B, e->lrc_pair + e->lrc_self, e->es_self, !cf.rd_only, cf.polarization, cf.polar_palmo,
		                                   d_per_bead, e->d_pisums.p);
	e->launches++;
	CK(cudaGetLastError());
	return MPMC_OK;
}

int mpmc_pi_potential(mpmc_engine *e, double *per_bead, double sums[4]) {
	CK(cudaSetDevice(e->dev));
	int rc = pi_sums_enqueue(e, nullptr);
	if (rc) return rc;
	// re-run the tiny assembly with the per-bead output placed right after the sums
	const mpmc_config &cf = e->cfg;
	k_pi_sums<<<1, 32, 0, e->stream>>>(e->d_result.p, e->B, e->lrc_pair + e->lrc_self, e->es_self, !cf.rd_only, cf.polarization, cf.polar_palmo,
	                                   e->d_pisums.p + 8, e->d_pisums.p);
	e->launches++;
	CK(cudaMemcpyAsync(e->h_pisums, e->d_pisums.p, sizeof(double) * (8 + 4 * (size_t)e->B), cudaMemcpyDeviceToHost, e->stream));
	std::vector<mpmc_energy_out> out(e->B);
	if ((rc = mpmc_energy_fetch(e, out.data()))) return rc;     // synchronises the stream (and keeps iterator_failed etc. current)
	for (int q = 0; q < 4; q++) sums[q] = e->h_pisums[q];
	if (per_bead) memcpy(per_bead, e->h_pisums + 8, sizeof(double) * 4 * e->B);
	return MPMC_OK;
}

int mpmc_nccl_get_unique_id(char id[128]) {
	int rc = load_nccl();
	if (rc) return rc;
	NK(g_nccl.GetUniqueId(id));
	return MPMC_OK;
}

int mpmc_nccl_init(mpmc_engine *e, const char id[128], int rank, int nranks) {
	int rc = load_nccl();
	if (rc) return rc;
	CK(cudaSetDevice(e->dev));
	if (rank < 0 || rank >= nranks) FAIL(MPMC_ERR_INVALID_SETTING, "nccl_init: bad rank %d of %d", rank, nranks);
	NcclId nid;
	memcpy(nid.internal, id, 128);
	NK(g_nccl.CommInitRank(&e->comm, nranks, nid, rank));
	e->rank = rank; e->nranks = nranks;
	// Peer-memory mailboxes for the fused all-reduce (k_pi_sums_xchg): one allocation per rank, its CUDA IPC handle all-gathered
	// through the communicator, every peer's mailbox mapped here.  If any step fails (no peer access, IPC unavailable) the engine
	// keeps using ncclAllReduce.  MPMC_PI_P2P=0 forces that path (A/B measurements).
	const char *env = getenv("MPMC_PI_P2P");
	if (nranks > 1 && nranks <= 32 && !(env && env[0] == '0')) {
		bool ok = true;
		const size_t bytes = sizeof(PiMailSlot) * 2 * nranks;
		cudaIpcMemHandle_t mine;
		char *d_h = nullptr;
		std::vector<cudaIpcMemHandle_t> all(nranks);
		ok = ok && cudaMalloc(&e->d_mbox, bytes) == cudaSuccess && cudaMemset(e->d_mbox, 0, bytes) == cudaSuccess;
		ok = ok && cudaIpcGetMemHandle(&mine, e->d_mbox) == cudaSuccess;
		ok = ok && cudaMalloc(&d_h, sizeof(mine) * (nranks + 1)) == cudaSuccess;
		ok = ok && cudaMemcpy(d_h + sizeof(mine) * nranks, &mine, sizeof(mine), cudaMemcpyHostToDevice) == cudaSuccess;
		// every rank must take part in the collective, whatever happened locally; a rank that failed sends a zero handle
		if (!ok && d_h) cudaMemset(d_h + sizeof(mine) * nranks, 0, sizeof(mine));
		if (d_h) {
			const int r = g_nccl.AllGather(d_h + sizeof(mine) * nranks, d_h, sizeof(mine), /*ncclChar*/ 0, e->comm, e->stream);
			ok = ok && r == 0 && cudaStreamSynchronize(e->stream) == cudaSuccess &&
			     cudaMemcpy(all.data(), d_h, sizeof(mine) * nranks, cudaMemcpyDeviceToHost) == cudaSuccess;
			cudaFree(d_h);
		}
		e->peer_mbox.assign(nranks, nullptr);
		for (int r = 0; ok && r < nranks; r++) {
			bool zero = true;
			for (size_t b = 0; b < sizeof(mine); b++) zero = zero && ((const char *)&all[r])[b] == 0;
			if (zero) { ok = false; break; }
			if (r == rank) { e->peer_mbox[r] = e->d_mbox; continue; }
			void *p = nullptr;
			ok = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
			e->peer_mbox[r] = (PiMailSlot *)p;
		}
		if (ok) ok = e->d_peers.ensure(nranks) == MPMC_OK && e->d_step.ensure(1) == MPMC_OK &&
		             cudaMemcpy(e->d_peers.p, e->peer_mbox.data(), sizeof(PiMailSlot *) * nranks, cudaMemcpyHostToDevice) == cudaSuccess &&
		             cudaMemset(e->d_step.p, 0, sizeof(long long)) == cudaSuccess;
		cudaGetLastError();
		// all ranks must agree: one more tiny collective on the verdict (sum of failures)
		double *d_v = nullptr;
		double v = ok ? 0.0 : 1.0;
		if (cudaMalloc(&d_v, sizeof(double)) == cudaSuccess) {
			cudaMemcpy(d_v, &v, sizeof(double), cudaMemcpyHostToDevice);
			if (g_nccl.AllReduce(d_v, d_v, 1, kNcclFloat64, kNcclSum, e->comm, e->stream) == 0 && cudaStreamSynchronize(e->stream) == cudaSuccess)
				cudaMemcpy(&v, d_v, sizeof(double), cudaMemcpyDeviceToHost);
			else v = 1.0;
			cudaFree(d_v);
		} else v = 1.0;
		e->p2p = (v == 0.0);
	}
	return MPMC_OK;
}

int mpmc_pi_potential_allreduce(mpmc_engine *e, int P_global, double means[4], double *potential) {
	CK(cudaSetDevice(e->dev));
	if (P_global < e->B) FAIL(MPMC_ERR_BEADS, "P_global (%d) smaller than the local bead count (%d)", P_global, e->B);
	int rc;
	if (!e->h_pisums) CK(cudaMallocHost(&e->h_pisums, sizeof(double) * (8 + 4 * (size_t)e->B)));
	// Steady state (same topology, no framework move pending, no per-kernel timing, no polarization loop with host decisions, the
	// structure-factor partials valid so that only the moved chunks are redone): the whole sweep — dirty-chunk list, structure
	// factor, pair sweep, finish (reductions + assembly + exchange), result copy — is one CUDA graph, captured on the second call and
	// replayed afterwards.  At 8 beads per GPU the sweep is ~100 us of kernels: separate launches would leave the GPU idle for a
	// third of that.
	const bool es = !e->cfg.rd_only;
	const bool graphable = e->pi_warm && !e->pi_graph_off && !e->timing && !e->cfg.polarization && !e->topo_dirty && !e->frozen_sk_dirty && (!es || e->sk_valid);
	// single GPU or peer-memory exchange (and no polarization loop): the sweep's last kernel posts sums + launch number into the pinned
	// result buffer itself; otherwise (NCCL fallback, polarizable beads) the result is copied after the last operation
	const bool direct = !e->cfg.polarization && (e->p2p || !e->comm);
	auto copy_back = [&]() { return direct ? cudaSuccess : cudaMemcpyAsync(e->h_pisums, e->d_pisums.p, sizeof(double) * 6, cudaMemcpyDeviceToHost, e->stream); };
	if (graphable && e->pi_graph) {
		// what run_structure_mobile does on the host when it is not replayed from a graph
		e->h_sk_dirty[0] = (int)e->sk_dirty.size();
		for (size_t i = 0; i < e->sk_dirty.size(); i++) e->h_sk_dirty[1 + i] = e->sk_dirty[i];
		e->sk_dirty.clear();
		CK(cudaGraphLaunch(e->pi_graph, e->stream));
		e->launches += e->pi_graph_launches;
	} else if (graphable) {
		const long long l0 = e->launches;
		const std::vector<int> dirty0 = e->sk_dirty;
		cudaGraph_t g = nullptr;
		CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
		rc = pi_sums_enqueue(e, nullptr, true, true);
		int nrc = 0;
		if (!rc && e->comm && !e->p2p) nrc = g_nccl.AllReduce(e->d_pisums.p, e->d_pisums.p, 4, kNcclFloat64, kNcclSum, e->comm, e->stream);
		cudaError_t ce = (!rc && !nrc) ? copy_back() : cudaSuccess;
		cudaError_t ee = cudaStreamEndCapture(e->stream, &g);
		if (rc || nrc || ce != cudaSuccess || ee != cudaSuccess || !g || cudaGraphInstantiate(&e->pi_graph, g, 0) != cudaSuccess) {
			cudaGetLastError();
			if (g) cudaGraphDestroy(g);
			e->pi_graph = nullptr; e->pi_graph_off = true;      // fall back to plain launches for the rest of this engine's life
			e->launches = l0;
			e->sk_valid = false;                                // nothing was executed: redo the structure factor in full
			if ((rc = pi_sums_enqueue(e, nullptr, true, true))) return rc;
			if (e->comm && !e->p2p) NK(g_nccl.AllReduce(e->d_pisums.p, e->d_pisums.p, 4, kNcclFloat64, kNcclSum, e->comm, e->stream));
			CK(copy_back());
		} else {
			cudaGraphDestroy(g);
			e->pi_graph_launches = e->launches - l0;
			// capture recorded the work without running it: the pinned dirty list the graph reads is the one staged during capture
			e->h_sk_dirty[0] = (int)dirty0.size();
			for (size_t i = 0; i < dirty0.size(); i++) e->h_sk_dirty[1 + i] = dirty0[i];
			CK(cudaGraphLaunch(e->pi_graph, e->stream));
		}
	} else {
		if ((rc = pi_sums_enqueue(e, nullptr, true, true))) return rc;
		if (e->comm && !e->p2p) NK(g_nccl.AllReduce(e->d_pisums.p, e->d_pisums.p, 4, kNcclFloat64, kNcclSum, e->comm, e->stream));
		CK(copy_back());
	}
	if (direct && !e->timing) {
		// wait for the launch number in host memory instead of a stream synchronisation (its wake-up costs ~10 us of a 100 us sweep);
		// if it does not come soon — a long exchange wait, a fault — fall back to the synchronisation, which also reports errors
		const double want = (double)(++e->pi_posted);
		volatile double *hp = e->h_pisums;
		long spins = 0;
		while (hp[7] != want && ++spins < 20000000L) { }
		if (hp[7] != want) { int _rc = sync_stream(e); if (_rc) return _rc; if (hp[7] != want) FAIL(MPMC_ERR_INTERNAL, "path-integral sweep finished without posting its result"); }
		e->stage_busy[0] = e->stage_busy[1] = false;        // everything staged before this sweep has been consumed
	} else {
		if (direct) ++e->pi_posted;
		int _rc = sync_stream(e); if (_rc) return _rc;
	}
	e->enqueued = false;
	e->pi_warm = true;
	if (e->timing) collect_timing(e);
	if (e->h_pisums[5] != 0.0) {
		FAIL(MPMC_ERR_CUDA, "path-integral exchange failed: a peer rank did not deliver its bead sums within the time limit (MPMC_PI_XCHG_TIMEOUT_S) or has stopped; "
		                    "every rank of the group reports this error");
	}
	if (e->gs_ran) {                                          // a polarizable bead system solved with the Gauss-Seidel pipeline
		e->gs_ran = false;
		if (*e->h_gs_abort) { *e->h_gs_abort = 0; e->gs_gen = 0; FAIL(MPMC_ERR_CUDA, "the Gauss-Seidel pipeline timed out waiting for its own CTAs"); }
	}
	for (int q = 0; q < 4; q++) means[q] = e->h_pisums[q] / P_global;          // PathIntegral.cpp:798-801
	if (potential) *potential = means[0] + means[1] + means[3] + means[2];     // :803-804
	return MPMC_OK;
}

int mpmc_pi_collective(mpmc_engine *e) { return e->p2p ? 2 : (e->comm ? 1 : 0); }

int mpmc_pi_chain_allreduce(mpmc_engine *e, double *chain_mass_len2) {
	CK(cudaSetDevice(e->dev));
	const int nmol = (int)e->mol_start.size() - 1, B = e->B;
	if (nmol < 1) FAIL(MPMC_ERR_NO_MOLECULES, "pi_chain: no sites uploaded");
	int rc;
	if ((rc = e->d_com.ensure((size_t)B * nmol * 3)) || (rc = e->d_mol_mass.ensure(nmol)) || (rc = e->d_chain.ensure(nmol)) ||
	    (rc = e->d_firstcom.ensure((size_t)e->nranks * nmol * 3)) || (rc = e->d_pisums.ensure(8 + 4 * (size_t)B))) return rc;
	if (!e->h_pisums) CK(cudaMallocHost(&e->h_pisums, sizeof(double) * (8 + 4 * (size_t)B)));
	k_mol_com<<<(nmol * B + 127) / 128, 128, 0, e->stream>>>(e->d_posq.p, e->cap, e->d_mass.p, e->d_mol_start.p, nmol, B, e->d_com.p, e->d_mol_mass.p);
	k_chain_len2<<<(nmol + 127) / 128, 128, 0, e->stream>>>(e->d_com.p, e->d_mol_mass.p, e->d_mol_mobile.p, nmol, B, e->nranks == 1, e->d_chain.p);
	e->launches += 2;
	if (e->nranks > 1) {
		// the link from my last bead to the first bead of the next rank: all-gather every rank's first-bead COMs
		NK(g_nccl.AllGather(e->d_com.p, e->d_firstcom.p, (size_t)nmol * 3, kNcclFloat64, e->comm, e->stream));
		k_chain_boundary<<<(nmol + 127) / 128, 128, 0, e->stream>>>(e->d_com.p + (size_t)(B - 1) * nmol * 3,
		                                                          e->d_firstcom.p + (size_t)((e->rank + 1) % e->nranks) * nmol * 3, e->d_mol_mass.p,
		                                                          e->d_mol_mobile.p, nmol, e->d_chain.p);
		e->launches++;
	}
	k_sum_array<<<1, 256, 0, e->stream>>>(e->d_chain.p, nmol, e->d_pisums.p + 4);
	e->launches++;
	if (e->comm) NK(g_nccl.AllReduce(e->d_pisums.p + 4, e->d_pisums.p + 4, 1, kNcclFloat64, kNcclSum, e->comm, e->stream));
	CK(cudaMemcpyAsync(e->h_pisums + 4, e->d_pisums.p + 4, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
	{ int _rc = sync_stream(e); if (_rc) return _rc; }
	*chain_mass_len2 = e->h_pisums[4];
	return MPMC_OK;
}

int mpmc_pi_chain(mpmc_engine *e, int closed, double *chain_mass_len2, double *com, double *mol_mass, int *n_mol) {
	CK(cudaSetDevice(e->dev));
	const int nmol = (int)e->mol_start.size() - 1, B = e->B;
	if (nmol < 1) FAIL(MPMC_ERR_NO_MOLECULES, "pi_chain: no sites uploaded");
	int rc;
	if ((rc = e->d_com.ensure((size_t)B * nmol * 3)) || (rc = e->d_mol_mass.ensure(nmol)) || (rc = e->d_chain.ensure(nmol))) return rc;
	k_mol_com<<<(nmol * B + 127) / 128, 128, 0, e->stream>>>(e->d_posq.p, e->cap, e->d_mass.p, e->d_mol_start.p, nmol, B, e->d_com.p, e->d_mol_mass.p);
	k_chain_len2<<<(nmol + 127) / 128, 128, 0, e->stream>>>(e->d_com.p, e->d_mol_mass.p, e->d_mol_mobile.p, nmol, B, closed, e->d_chain.p);
	e->launches += 2;
	std::vector<double> per(nmol);
	CK(cudaMemcpyAsync(per.data(), e->d_chain.p, sizeof(double) * nmol, cudaMemcpyDeviceToHost, e->stream));
	if (com) CK(cudaMemcpyAsync(com, e->d_com.p, sizeof(double) * (size_t)B * nmol * 3, cudaMemcpyDeviceToHost, e->stream));
	if (mol_mass) CK(cudaMemcpyAsync(mol_mass, e->d_mol_mass.p, sizeof(double) * nmol, cudaMemcpyDeviceToHost, e->stream));
	{ int _rc = sync_stream(e); if (_rc) return _rc; }
	double s = 0;
	for (int m = 0; m < nmol; m++) s += per[m];      // PathIntegral.cpp:880-901, molecule order
	*chain_mass_len2 = s;
	if (n_mol) *n_mol = nmol;
	return MPMC_OK;
}

// developer hook (not part of the drop-in surface): cycle stamps of the Gauss-Seidel pipeline's last launch
int mpmc_debug_gs_profile(mpmc_engine *e, int enable, long long *out, int max_blocks, int *nblk) {
	CK(cudaSetDevice(e->dev));
	e->gs_prof_enabled = enable != 0;
	{ const int dbg = enable & ~1; CK(cudaMemcpyToSymbol(g_gs_debug, &dbg, sizeof(int))); }
	if (nblk) *nblk = e->gs_grid;
	if (out && e->d_gsprof.p && e->gs_prof_nblk) {
		{ int _rc = sync_stream(e); if (_rc) return _rc; }
		const int nb = std::min(max_blocks, e->gs_prof_nblk);
		CK(cudaMemcpy(out, e->d_gsprof.p, sizeof(long long) * 8 * nb, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(out + (size_t)8 * max_blocks, e->d_gsprof.p + (size_t)8 * e->gs_prof_nblk, sizeof(long long) * 8 * nb, cudaMemcpyDeviceToHost));
		// (optional third part, when the caller's buffer has room for it: per-chunk hand-over stamps, 16 chunks per block)
		if (enable & 0x100) CK(cudaMemcpy(out + (size_t)16 * max_blocks, e->d_gsprof.p + (size_t)16 * e->gs_prof_nblk, sizeof(long long) * 16 * nb, cudaMemcpyDeviceToHost));
		if (nblk) *nblk = nb;
	}
	return MPMC_OK;
}

// developer hook: per-warp timeline of the pair sweep.  enable -> the next evaluations record it; out != NULL -> copy the last one
// back: 4 words per warp { entry, first sites loaded, last item summed (globaltimer ns), items done | SM << 32 }
int mpmc_debug_pair_profile(mpmc_engine *e, int enable, long long *out, int max_warps, int *nwarps) {
	CK(cudaSetDevice(e->dev));
	e->pair_prof_enabled = enable != 0;
	if (nwarps) *nwarps = e->pair_prof_warps;
	if (out && e->d_pairprof.p && e->pair_prof_warps) {
		{ int _rc = sync_stream(e); if (_rc) return _rc; }
		const int nw = std::min(max_warps, e->pair_prof_warps);
		CK(cudaMemcpy(out, e->d_pairprof.p, sizeof(long long) * 4 * nw, cudaMemcpyDeviceToHost));
		if (nwarps) *nwarps = nw;
	}
	return MPMC_OK;
}

// developer / bench hook: treat sites [first, first + count) as moved (their structure-factor chunks are recomputed by the next
// evaluation) without sending coordinates — what a device-resident timing loop uses so that no part of a sweep is skipped
int mpmc_debug_mark_moved(mpmc_engine *e, int first, int count) {
	if (first < 0 || count < 0 || first + count > e->n) FAIL(MPMC_ERR_INVALID_INPUT, "mark_moved: range out of bounds");
	mark_moved(e, first, count);
	return MPMC_OK;
}

// developer / test hooks for the host-side numerics (no device needed): the radial tables and the r^2 cutoff thresholds
int mpmc_debug_radial_table(int kind, double param, double u_lo, double u_hi, const double *u, int n, double *out0, double *out1) {
	RadialTable t;
	const long double a = param, osp = 0.5641895835477562869480794515607725858440506293289988L;
	if (kind == 0) t.build(1, u_lo, u_hi, [&](long double x, long double *o) { const long double r = sqrtl(x); o[0] = erfcl(a * r) / r; });
	else if (kind == 1) t.build(2, u_lo, u_hi, [&](long double x, long double *o) {
		const long double r = sqrtl(x), g = 2.0L * a * osp * expl(-a * a * x) * r;
		o[0] = (g + erfcl(a * r)) / (x * r); o[1] = (g - erfl(a * r)) / (x * r); }, kTabShiftCoarse);
	else FAIL(MPMC_ERR_INVALID_INPUT, "unknown table kind %d", kind);
	for (int i = 0; i < n; i++) {
		if (!(u[i] >= t.u_lo && u[i] < t.u_hi)) FAIL(MPMC_ERR_INVALID_INPUT, "u[%d] = %g outside the table [%g, %g)", i, u[i], t.u_lo, t.u_hi);
		out0[i] = t.eval(0, u[i]);
		if (kind == 1 && out1) out1[i] = t.eval(1, u[i]);
	}
	return MPMC_OK;
}

int mpmc_debug_cutoff_thresholds(double cutoff, double out[2]) {
	const volatile double rc = cutoff;
	const double top = 4.0 * cutoff * cutoff + 1.0;
	out[0] = largest_true([&](double x) { volatile double r = std::sqrt(x); volatile double d = r - kSmallDr; return d < rc; }, top);
	out[1] = largest_true([&](double x) { volatile double r = std::sqrt(x); return !(r > rc); }, top);
	return MPMC_OK;
}

int mpmc_set_timing(mpmc_engine *e, int on) {
	CK(cudaSetDevice(e->dev));
	{ int _rc = sync_stream(e); if (_rc) return _rc; }
	e->timing = on != 0;
	e->ev_used = 0;
	for (int i = 0; i < MPMC_NUM_KERNEL_CLASSES; i++) { e->t_ms[i] = 0; e->t_count[i] = 0; }
	return MPMC_OK;
}

int mpmc_get_timing(mpmc_engine *e, double ms[MPMC_NUM_KERNEL_CLASSES], long long count[MPMC_NUM_KERNEL_CLASSES]) {
	for (int i = 0; i < MPMC_NUM_KERNEL_CLASSES; i++) { ms[i] = e->t_ms[i]; count[i] = e->t_count[i]; }
	return MPMC_OK;
}

void *mpmc_stream(mpmc_engine *e) { return (void *)e->stream; }
long long mpmc_kernel_launches(mpmc_engine *e) { return e->launches; }

int mpmc_probe_fp64_peak(int device, double *tflops, double *sm_clock_mhz_guess) {
	CK(cudaSetDevice(device));
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, device));
	const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
	double *d = nullptr;
	CK(cudaMalloc(&d, sizeof(double) * blocks * threads));
	cudaEvent_t a, b;
	CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	k_fp64_probe<<<blocks, threads>>>(d, iters);   // warm-up
	CK(cudaDeviceSynchronize());
	float best = 1e30f;
	for (int rep = 0; rep < 5; rep++) {
		CK(cudaEventRecord(a));
		k_fp64_probe<<<blocks, threads>>>(d, iters);
		CK(cudaEventRecord(b));
		CK(cudaEventSynchronize(b));
		float ms = 0;
		CK(cudaEventElapsedTime(&ms, a, b));
		best = std::min(best, ms);
	}
	const double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
	*tflops = flops / (best * 1e-3) / 1e12;
	if (sm_clock_mhz_guess) *sm_clock_mhz_guess = *tflops * 1e12 / (2.0 * 64.0 * prop.multiProcessorCount) / 1e6;   // if 64 FP64 lanes per SM
	cudaFree(d); cudaEventDestroy(a); cudaEventDestroy(b);
	return MPMC_OK;
}

} // extern "C"
