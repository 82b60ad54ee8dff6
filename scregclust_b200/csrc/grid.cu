// Whole-grid fit: the control flow of scregclust() (R/scregclust.R:1222-1912) for a penalization grid,
// on top of the batched per-cycle entry points of a session.  Host code; the numerics are in session.cu.
//
// Every penalization value that is still iterating is advanced by the same batched device call (the fits
// are independent: R/scregclust.R:1223,1245).  In a multi-GPU job every rank runs this loop on identical
// inputs; the device calls are collective (session.cu) and return identical outputs on every rank, so the
// ranks take identical decisions without exchanging anything here.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <thread>

#include "session.cuh"

namespace scrc {

// ---------------------------------------------------------------------------------------------- backend
struct GridBackend {
    int64_t P = 0, T = 0;
    int K = 0;
    virtual ~GridBackend() {}
    virtual void fit_cycle(int F, const double* lambda, const int* k_in, const scrc_options& opt, bool skip_alloc,
                           const int* hint, const scrc_prior* prior, const scrc_quick* quick, scrc_cycle_out* out) = 0;
    virtual void importance(int F, const int* k_final, const unsigned char* models, const double* signs,
                            const scrc_options& opt, double* r2_module, double* importance, double* r2_removed,
                            int64_t cap) = 0;
    // coefficient matrices of fit f of the last skip_allocation cycle: packed size known from n_selected
    virtual void coeffs_fit(int f, int64_t ndoubles, int nsel, std::vector<double>& host, DevBuf<double>& dev,
                            std::vector<int>& sel) = 0;
    virtual int device() const { return -1; }
    virtual void request_cancel() {}
};

struct CtxBackend : GridBackend {
    Ctx& c;
    explicit CtxBackend(Ctx& ctx) : c(ctx) { P = c.P; T = c.T; K = c.K; }
    void fit_cycle(int F, const double* lambda, const int* k_in, const scrc_options& opt, bool skip_alloc, const int* hint,
                   const scrc_prior* prior, const scrc_quick* quick, scrc_cycle_out* out) override {
        if (hint) c.cost_hint.assign(hint, hint + (size_t)F * K);
        c.fit_cycle(F, lambda, k_in, opt, skip_alloc, out, prior, quick);
    }
    void importance(int F, const int* k_final, const unsigned char* models, const double* signs, const scrc_options& opt,
                    double* r2_module, double* imp, double* r2_removed, int64_t cap) override {
        c.importance(F, k_final, models, signs, opt, r2_module, imp, r2_removed, cap);
    }
    void coeffs_fit(int f, int64_t ndoubles, int nsel, std::vector<double>&, DevBuf<double>& dev, std::vector<int>& sel) override {
        // stays on the device (packed) until the caller asks for it: no staging copy on the host
        sel.resize(nsel);
        c.get_coeffs_fit(f, nullptr, sel.data());
        AllocScope as(c.dev.stream);           // pooled, stream-ordered allocation (freed with the result handle)
        dev.alloc((size_t)ndoubles);
        c.pack_coeffs_fit(f, dev.p);
    }
    int device() const override { return c.dev.id; }
    void request_cancel() override { c.cancel.store(1, std::memory_order_relaxed); }
};

struct CallbackBackend : GridBackend {
    scrc_grid_backend cb;
    explicit CallbackBackend(const scrc_grid_backend& b) : cb(b) { P = b.P; T = b.T; K = b.K; }
    static void chk(int rc, const char* what) {
        if (rc != 0) throw Error(rc, std::string("scrc_grid_fit_custom: the ") + what + " callback failed");
    }
    void fit_cycle(int F, const double* lambda, const int* k_in, const scrc_options& opt, bool skip_alloc, const int* hint,
                   const scrc_prior* prior, const scrc_quick* quick, scrc_cycle_out* out) override {
        chk(cb.fit_cycle(cb.user, F, lambda, k_in, &opt, skip_alloc ? 1 : 0, hint, prior, quick, out), "fit_cycle");
    }
    void importance(int F, const int* k_final, const unsigned char* models, const double* signs, const scrc_options& opt,
                    double* r2_module, double* imp, double* r2_removed, int64_t cap) override {
        if (!cb.importance) throw Error(SCRC_ERR_ARG, "scrc_grid_fit_custom: compute_predictive_r2 needs an importance callback");
        chk(cb.importance(cb.user, F, k_final, models, signs, &opt, r2_module, imp, r2_removed, cap), "importance");
    }
    void coeffs_fit(int f, int64_t ndoubles, int nsel, std::vector<double>& host, DevBuf<double>&, std::vector<int>& sel) override {
        if (!cb.coeffs_fit) throw Error(SCRC_ERR_ARG, "scrc_grid_fit_custom: keep_coeffs needs a coeffs_fit callback");
        host.assign((size_t)ndoubles, 0.0);
        sel.assign(nsel, 0);
        chk(cb.coeffs_fit(cb.user, f, host.data(), sel.data()), "coeffs_fit");
    }
};

// ---------------------------------------------------------------------------------------------- result
struct CycleRecord {
    std::vector<int> admm_iterations, n_selected;      // K
    std::vector<unsigned char> models;                  // K x P
    std::vector<int> votes, nnls_iterations;            // diagnostics only
};
struct FinalConfig {
    std::vector<int> k_final, best_r2_idx, n_selected, sel;
    std::vector<unsigned char> models;
    std::vector<double> weights, signs, r2_test, resid_var, best_r2, r2_module, importance, r2_removed;
    std::vector<double> coeffs_host;
    DevBuf<double> coeffs_dev;
    int64_t coeffs_n = 0;
    bool have_coeffs = false;
};
struct PenaltyResult {
    double penalization = 0.0;
    bool converged = false;
    int n_cycles_run = 0, final_cycle_length = 1;
    std::vector<std::vector<int>> k_history;
    std::vector<CycleRecord> records;
    std::vector<FinalConfig> finals;
};
}  // namespace scrc

struct scrc_grid_result {
    std::vector<scrc::PenaltyResult> pens;
    scrc_grid_info info{};
    int device = -1;
};

namespace scrc {

struct PenState {
    std::vector<int> k;
    int n_k = 2, cycle = 0;
    bool last_cycle = false;
    std::vector<int> last_iters;     // ADMM iterations of the previous cycle per module (cost hint)
};

static double wall_now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static void grid_loop(GridBackend& be, const scrc_grid_params& gp, scrc_grid_result& res) {
    const bool dbg_t = getenv("SCRC_DEBUG_TIMING") != nullptr;
    const double t_loop0 = wall_now();
    double t_cycle = 0.0, t_final = 0.0, t_imp = 0.0, t_coef = 0.0;
    const int L = gp.n_penalties, K = be.K;
    const int64_t P = be.P, T = be.T;
    if (L <= 0 || !gp.penalization || !gp.initial_target_modules) throw Error(SCRC_ERR_ARG, "scrc_grid_fit: empty grid or null pointer");
    if (gp.n_cycles < 0 || gp.min_module_size < 0) throw Error(SCRC_ERR_ARG, "scrc_grid_fit: negative n_cycles / min_module_size");
    if (gp.quick_mode && !(gp.quick_mode_percent >= 0.0 && gp.quick_mode_percent < 1.0))
        throw Error(SCRC_ERR_ARG, "quick_mode_percent needs to be numeric in [0, 1).");
    if (gp.quick_mode && gp.prior_weight != 0.0 && gp.prior_off)
        throw Error(SCRC_ERR_UNSUPPORTED, "scrc_grid_fit: quick mode together with a network prior is not supported");
    const bool use_prior = gp.prior_weight != 0.0 && gp.prior_off != nullptr;
    const bool aggregate = gp.allocate_per_obs == 0;
    if (aggregate && gp.quick_mode)                                                       // R/scregclust.R:807-817
        throw Error(SCRC_ERR_ARG, "quick_mode only available when allocate_per_obs is TRUE.");
    const bool diag = gp.keep_diagnostics != 0;
    res.pens.clear();
    res.pens.resize(L);
    res.info = scrc_grid_info{};
    res.info.n_penalties = L; res.info.P = P; res.info.T = T; res.info.K = K;
    res.device = be.device();
    std::vector<PenState> st(L);
    for (int i = 0; i < L; ++i) {
        st[i].k.assign(gp.initial_target_modules, gp.initial_target_modules + T);
        res.pens[i].penalization = gp.penalization[i];
        res.pens[i].k_history.push_back(st[i].k);
        st[i].last_iters.assign(K, 0);
    }
    std::vector<int> active(L), pending_final;
    for (int i = 0; i < L; ++i) active[i] = i;

    std::vector<double> lam;
    std::vector<int> ks, hint, k_new, admm_it, nsel, votes, nnls_it, orders;
    std::vector<unsigned char> models;
    std::vector<double> best_r2;
    std::vector<int> identity;
    long long sweeps = 0;

    while (!active.empty()) {
        // R/scregclust.R:1269-1284: cycle counter, forced last cycle after n_cycles
        std::vector<int> run, fin;
        for (int i : active) {
            PenState& s = st[i];
            s.cycle += 1;
            if (s.cycle == gp.n_cycles + 1) s.last_cycle = true;
            res.pens[i].n_cycles_run = s.cycle;          // the last cycle (re-fit only) counts, R/scregclust.R:1269
            (s.last_cycle ? fin : run).push_back(i);
        }
        if (!run.empty()) {
            const int F = (int)run.size();
            lam.resize(F); ks.resize((size_t)F * T); hint.resize((size_t)F * K);
            k_new.assign((size_t)F * T, 0); best_r2.assign((size_t)F * T, 0.0);
            models.assign((size_t)F * K * P, 0); admm_it.assign((size_t)F * K, 0); nsel.assign((size_t)F * K, 0);
            for (int f = 0; f < F; ++f) {
                const PenState& s = st[run[f]];
                lam[f] = gp.penalization[run[f]];
                std::copy(s.k.begin(), s.k.end(), ks.begin() + (size_t)f * T);
                std::copy(s.last_iters.begin(), s.last_iters.end(), hint.begin() + (size_t)f * K);
            }
            scrc_cycle_out co{};
            co.k_new = k_new.data(); co.best_r2 = best_r2.data(); co.models = models.data();
            co.admm_iterations = admm_it.data(); co.n_selected = nsel.data(); co.admm_sweeps = &sweeps;
            if (diag) {
                votes.assign((size_t)F * T * K, 0); nnls_it.assign((size_t)F * K * T, 0);
                co.votes = votes.data(); co.nnls_iterations = nnls_it.data();
            }
            scrc_prior pr{};
            if (use_prior) {
                orders.resize((size_t)F * T);
                for (int f = 0; f < F; ++f) {
                    const int* o = gp.update_order ? gp.update_order(gp.update_order_user, run[f], st[run[f]].cycle) : nullptr;
                    if (o) std::copy(o, o + T, orders.begin() + (size_t)f * T);
                    else for (int64_t t = 0; t < T; ++t) orders[(size_t)f * T + t] = (int)t;
                }
                pr.prior_off = gp.prior_off; pr.prior_idx = gp.prior_idx; pr.baseline = gp.prior_baseline;
                pr.weight = gp.prior_weight; pr.update_order = orders.data();
            }
            if (aggregate) {   // allocate_per_obs = FALSE: the prior (if any) enters through the same structure
                pr.aggregate = 1; pr.baseline = gp.prior_baseline; pr.weight = gp.prior_weight;
            }
            sweeps = 0;
            if (gp.cancel && *gp.cancel) be.request_cancel();
            const double tc0 = wall_now();
            // quick mode from cycle two on (R/scregclust.R:1444); in a fit cycle R's `idx` is the configuration itself
            scrc_quick qk{ks.data(), gp.quick_mode_percent};
            const bool quick_now = gp.quick_mode && st[run[0]].cycle != 1;
            be.fit_cycle(F, lam.data(), ks.data(), gp.opt, false, hint.data(), (use_prior || aggregate) ? &pr : nullptr,
                         quick_now ? &qk : nullptr, &co);
            t_cycle += wall_now() - tc0;
            res.info.admm_sweeps += sweeps; res.info.device_calls += 1;
            for (int f = 0; f < F; ++f) {
                const int i = run[f];
                PenState& s = st[i];
                PenaltyResult& pr_ = res.pens[i];
                CycleRecord rec;
                rec.admm_iterations.assign(admm_it.begin() + (size_t)f * K, admm_it.begin() + (size_t)(f + 1) * K);
                rec.n_selected.assign(nsel.begin() + (size_t)f * K, nsel.begin() + (size_t)(f + 1) * K);
                rec.models.assign(models.begin() + (size_t)f * K * P, models.begin() + (size_t)(f + 1) * K * P);
                if (diag) {
                    rec.votes.assign(votes.begin() + (size_t)f * T * K, votes.begin() + (size_t)(f + 1) * T * K);
                    rec.nnls_iterations.assign(nnls_it.begin() + (size_t)f * K * T, nnls_it.begin() + (size_t)(f + 1) * K * T);
                }
                for (int j = 0; j < K; ++j) { res.info.admm_problem_iterations += rec.admm_iterations[j]; s.last_iters[j] = rec.admm_iterations[j]; }
                pr_.records.push_back(std::move(rec));
                // noise module and minimum module size (R/scregclust.R:1731-1739)
                std::vector<int> idx(k_new.begin() + (size_t)f * T, k_new.begin() + (size_t)(f + 1) * T);
                const double* br = best_r2.data() + (size_t)f * T;
                for (int64_t t = 0; t < T; ++t) if (br[t] < gp.noise_threshold) idx[t] = -1;
                std::vector<int> sizes(K + 1, 0);
                for (int64_t t = 0; t < T; ++t) if (idx[t] >= 1 && idx[t] <= K) ++sizes[idx[t]];
                bool any_small = false;
                for (int j = 1; j <= K; ++j) any_small |= sizes[j] < gp.min_module_size;
                if (any_small)
                    for (int64_t t = 0; t < T; ++t)
                        if (idx[t] >= 1 && idx[t] <= K && sizes[idx[t]] < gp.min_module_size) idx[t] = -1;
                // convergence / limit cycle (R/scregclust.R:1763-1802): first earlier configuration equal to the new one
                int match = 0;
                for (size_t h = 0; h < pr_.k_history.size(); ++h)
                    if (pr_.k_history[h] == idx) { match = (int)h + 1; break; }
                pr_.k_history.push_back(idx);
                if (match) {
                    pr_.converged = true;
                    if (s.n_k - match != 1) pr_.final_cycle_length = s.n_k - match;
                    s.last_cycle = true;
                }
                s.n_k += 1;
                s.k = std::move(idx);
            }
        }
        // Penalization values entering their last cycle only need the re-fit of their final configuration(s)
        // (R/scregclust.R:1307-1311, Step 2b skipped) and the post-processing.  Nothing depends on those results,
        // so they are deferred and batched into as few device calls as possible.
        pending_final.insert(pending_final.end(), fin.begin(), fin.end());
        std::vector<int> still;
        for (int i : active) if (std::find(fin.begin(), fin.end(), i) == fin.end()) still.push_back(i);
        active.swap(still);
    }

    auto finalize = [&](const std::vector<int>& batch) {
        std::vector<int> owner;
        std::vector<const std::vector<int>*> kf;
        for (int i : batch) {
            const PenaltyResult& pr_ = res.pens[i];
            for (int m = 1; m <= pr_.final_cycle_length; ++m) {                         // R/scregclust.R:1307-1311
                owner.push_back(i);
                kf.push_back(&pr_.k_history[pr_.k_history.size() - m]);
            }
        }
        const int F = (int)owner.size();
        lam.resize(F); ks.resize((size_t)F * T); hint.assign((size_t)F * K, 0);
        std::vector<int> latest;
        if (gp.quick_mode) latest.resize((size_t)F * T);
        for (int f = 0; f < F; ++f) {
            lam[f] = gp.penalization[owner[f]];
            std::copy(kf[f]->begin(), kf[f]->end(), ks.begin() + (size_t)f * T);
            // R's `idx` stays the LATEST allocation while older configurations of a limit cycle are re-fitted
            if (gp.quick_mode) std::copy(st[owner[f]].k.begin(), st[owner[f]].k.end(), latest.begin() + (size_t)f * T);
            std::copy(st[owner[f]].last_iters.begin(), st[owner[f]].last_iters.end(), hint.begin() + (size_t)f * K);
        }
        const size_t fkp = (size_t)F * K * P, fkt = (size_t)F * K * T, ft = (size_t)F * T;
        std::vector<unsigned char> md(fkp);
        std::vector<double> wt(fkp), sg(fkp), r2(fkt), rv(fkt), br(ft);
        std::vector<int> bi(ft), ai((size_t)F * K), ns((size_t)F * K);
        scrc_cycle_out co{};
        co.models = md.data(); co.weights = wt.data(); co.signs = sg.data(); co.r2_test = r2.data(); co.resid_var = rv.data();
        co.best_r2 = br.data(); co.best_r2_idx = bi.data(); co.admm_iterations = ai.data(); co.n_selected = ns.data();
        co.admm_sweeps = &sweeps;
        if (diag) { nnls_it.assign(fkt, 0); co.nnls_iterations = nnls_it.data(); }
        sweeps = 0;
        if (gp.cancel && *gp.cancel) be.request_cancel();
        const double tf0 = wall_now();
        scrc_quick qk{latest.data(), gp.quick_mode_percent};
        // (a value that never ran a fit cycle -- n_cycles = 0 -- is still in "cycle 1": no subset)
        bool quick_fin = gp.quick_mode != 0;
        for (int i : batch) quick_fin &= st[i].cycle != 1;
        be.fit_cycle(F, lam.data(), ks.data(), gp.opt, true, hint.data(), nullptr, quick_fin ? &qk : nullptr, &co);
        t_final += wall_now() - tf0;
        res.info.admm_sweeps += sweeps; res.info.device_calls += 1;
        for (int f = 0; f < F; ++f) {
            PenaltyResult& pr_ = res.pens[owner[f]];
            CycleRecord rec;
            rec.admm_iterations.assign(ai.begin() + (size_t)f * K, ai.begin() + (size_t)(f + 1) * K);
            rec.n_selected.assign(ns.begin() + (size_t)f * K, ns.begin() + (size_t)(f + 1) * K);
            rec.models.assign(md.begin() + (size_t)f * K * P, md.begin() + (size_t)(f + 1) * K * P);
            if (diag) rec.nnls_iterations.assign(nnls_it.begin() + (size_t)f * K * T, nnls_it.begin() + (size_t)(f + 1) * K * T);
            for (int j = 0; j < K; ++j) res.info.admm_problem_iterations += rec.admm_iterations[j];
            pr_.records.push_back(std::move(rec));
            pr_.finals.emplace_back();
            FinalConfig& fc = pr_.finals.back();
            fc.k_final = *kf[f];
            fc.models.assign(md.begin() + (size_t)f * K * P, md.begin() + (size_t)(f + 1) * K * P);
            fc.weights.assign(wt.begin() + (size_t)f * K * P, wt.begin() + (size_t)(f + 1) * K * P);
            fc.signs.assign(sg.begin() + (size_t)f * K * P, sg.begin() + (size_t)(f + 1) * K * P);
            fc.r2_test.assign(r2.begin() + (size_t)f * K * T, r2.begin() + (size_t)(f + 1) * K * T);
            fc.resid_var.assign(rv.begin() + (size_t)f * K * T, rv.begin() + (size_t)(f + 1) * K * T);
            fc.best_r2.assign(br.begin() + (size_t)f * T, br.begin() + (size_t)(f + 1) * T);
            fc.best_r2_idx.assign(bi.begin() + (size_t)f * T, bi.begin() + (size_t)(f + 1) * T);
            fc.n_selected.assign(ns.begin() + (size_t)f * K, ns.begin() + (size_t)(f + 1) * K);
            if (gp.keep_coeffs) {
                int tot = 0;
                for (int j = 0; j < K; ++j) tot += fc.n_selected[j];
                fc.coeffs_n = (int64_t)tot * T;
                fc.have_coeffs = true;
                const double tk0 = wall_now();
                if (tot > 0) be.coeffs_fit(f, fc.coeffs_n, tot, fc.coeffs_host, fc.coeffs_dev, fc.sel);
                t_coef += wall_now() - tk0;
            }
        }
        if (gp.compute_predictive_r2) {                                                   // R/scregclust.R:1823-1912
            std::vector<double> r2m((size_t)F * K), imp(fkp);
            // capacity of the packed r2_removed blocks: sum over (fit, module) with targets and > 1 regulator
            std::vector<int64_t> block((size_t)F, 0);
            int64_t cap = 0;
            for (int f = 0; f < F; ++f) {
                std::vector<int> msz(K + 1, 0);
                for (int64_t t = 0; t < T; ++t) { const int j = (*kf[f])[t]; if (j >= 1 && j <= K) ++msz[j]; }
                for (int j = 0; j < K; ++j)
                    if (msz[j + 1] > 0 && ns[(size_t)f * K + j] > 1) block[f] += (int64_t)msz[j + 1] * (ns[(size_t)f * K + j] + 1);
                cap += block[f];
            }
            std::vector<double> rem((size_t)std::max<int64_t>(cap, 1));
            const double ti0 = wall_now();
            be.importance(F, ks.data(), md.data(), sg.data(), gp.opt, r2m.data(), imp.data(), rem.data(), cap);
            t_imp += wall_now() - ti0;
            res.info.device_calls += 1;
            int64_t off = 0;
            // finals were appended in the order of `owner`, so walk them in the same order
            std::vector<int> seen(L, 0);
            for (int f = 0; f < F; ++f) {
                PenaltyResult& pr_ = res.pens[owner[f]];
                FinalConfig& fc = pr_.finals[pr_.finals.size() - pr_.final_cycle_length + seen[owner[f]]++];
                fc.r2_module.assign(r2m.begin() + (size_t)f * K, r2m.begin() + (size_t)(f + 1) * K);
                fc.importance.assign(imp.begin() + (size_t)f * K * P, imp.begin() + (size_t)(f + 1) * K * P);
                fc.r2_removed.assign(rem.begin() + off, rem.begin() + off + block[f]);
                off += block[f];
            }
        }
        for (int i : batch) { res.info.final_refits += res.pens[i].final_cycle_length; }
    };
    // at most ~2 L fits per call (bounds the state matrices of the re-fit)
    std::vector<int> batch;
    int width = 0;
    for (int i : pending_final) {
        const int w = res.pens[i].final_cycle_length;
        if (!batch.empty() && width + w > 2 * L) { finalize(batch); batch.clear(); width = 0; }
        batch.push_back(i); width += w;
    }
    if (!batch.empty()) finalize(batch);
    for (int i = 0; i < L; ++i) res.info.fit_cycles += res.pens[i].n_cycles_run - 1;
    if (dbg_t) {
        const double tot = wall_now() - t_loop0;
        fprintf(stderr, "[scrc grid] total %.1f ms: fit cycles %.1f, final re-fit %.1f, coefficient packing %.1f, importance %.1f, "
                        "loop logic %.1f\n", tot * 1e3, t_cycle * 1e3, t_final * 1e3, t_coef * 1e3, t_imp * 1e3,
                (tot - t_cycle - t_final - t_coef - t_imp) * 1e3);
    }
}

}  // namespace scrc

// ================================================================================================ C ABI
using namespace scrc;

extern "C" void scrc_set_last_error_(const char* msg);   // capi.cu

template <class F>
static int grid_guarded(F&& f) {
    try { f(); return SCRC_OK; }
    catch (const Error& e) { scrc_set_last_error_(e.what()); return e.code; }
    catch (const std::bad_alloc&) { scrc_set_last_error_("host allocation failed"); return SCRC_ERR_INTERNAL; }
    catch (const std::exception& e) { scrc_set_last_error_(e.what()); return SCRC_ERR_INTERNAL; }
    catch (...) { scrc_set_last_error_("unknown failure"); return SCRC_ERR_INTERNAL; }
}

static void check_params(const scrc_grid_params* p) {
    if (!p) throw Error(SCRC_ERR_ARG, "scrc_grid_fit: null parameters");
    const scrc_options& o = p->opt;
    if (o.max_optim_iter <= 0 || o.n_update <= 0 || !(o.rho0 > 0.0) || o.alpha0 <= 0.0 || o.alpha0 > 2.0)
        throw Error(SCRC_ERR_ARG, "scrc_grid_fit: invalid options");
}

extern "C" {

void scrc_grid_default_params(scrc_grid_params* p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->n_cycles = 50; p->noise_threshold = 0.025; p->min_module_size = 0;
    scrc_default_options(&p->opt);
    p->compute_predictive_r2 = 1; p->keep_coeffs = 1; p->keep_diagnostics = 0;
    p->prior_baseline = 1e-6; p->prior_weight = 0.0;
    p->quick_mode = 0; p->quick_mode_percent = 0.1; p->allocate_per_obs = 1;
}

int scrc_grid_fit(scrc_ctx* ctx, const scrc_grid_params* p, scrc_grid_result** out) {
    if (!ctx || !out) { scrc_set_last_error_("scrc_grid_fit: null pointer"); return SCRC_ERR_ARG; }
    *out = nullptr;
    return grid_guarded([&] {
        check_params(p);
        std::unique_ptr<scrc_grid_result> r(new scrc_grid_result());
        CtxBackend be(ctx->impl);
        grid_loop(be, *p, *r);
        *out = r.release();
    });
}

int scrc_grid_fit_custom(const scrc_grid_backend* backend, const scrc_grid_params* p, scrc_grid_result** out) {
    if (!backend || !out || !backend->fit_cycle) { scrc_set_last_error_("scrc_grid_fit_custom: null pointer"); return SCRC_ERR_ARG; }
    *out = nullptr;
    return grid_guarded([&] {
        check_params(p);
        if (backend->P <= 0 || backend->T <= 0 || backend->K <= 0) throw Error(SCRC_ERR_ARG, "scrc_grid_fit_custom: non-positive dimension");
        std::unique_ptr<scrc_grid_result> r(new scrc_grid_result());
        CallbackBackend be(*backend);
        grid_loop(be, *p, *r);
        *out = r.release();
    });
}

int scrc_grid_fit_multi(const int* devices, int n_devices, const double* z1_reg, const double* z2_reg,
                        const double* z1_target, const double* z2_target, const double* z1_target_sds, int64_t n1,
                        int64_t n2, int64_t P, int64_t T, int K, const scrc_grid_params* p, scrc_grid_result** out) {
    if (!devices || n_devices <= 0 || !out || !z1_reg || !z2_reg || !z1_target || !z2_target || !z1_target_sds) {
        scrc_set_last_error_("scrc_grid_fit_multi: null pointer or no device");
        return SCRC_ERR_ARG;
    }
    *out = nullptr;
    return grid_guarded([&] {
        check_params(p);
        char uid[128];
        if (n_devices > 1) comm_unique_id(uid);
        std::vector<std::unique_ptr<scrc_grid_result>> results(n_devices);
        std::vector<int> codes(n_devices, SCRC_OK);
        std::vector<std::string> msgs(n_devices);
        auto worker = [&](int r) {
            try {
                Comm comm;
                if (n_devices > 1) comm.init(uid, n_devices, r, devices[r]);
                std::unique_ptr<Ctx> ctx(new Ctx());
                ctx->create(devices[r], z1_reg, z2_reg, z1_target, z2_target, z1_target_sds, n1, n2, P, T, K,
                            n_devices > 1 ? &comm : nullptr);
                std::unique_ptr<scrc_grid_result> res(new scrc_grid_result());
                CtxBackend be(*ctx);
                grid_loop(be, *p, *res);
                SCRC_CUDA(cudaSetDevice(devices[r]));
                ctx.reset();                     // the session goes before its communicator
                if (r == 0) results[0] = std::move(res);
            } catch (const Error& e) { codes[r] = e.code; msgs[r] = e.what(); }
            catch (const std::exception& e) { codes[r] = SCRC_ERR_INTERNAL; msgs[r] = e.what(); }
        };
        std::vector<std::thread> th;
        for (int r = 1; r < n_devices; ++r) th.emplace_back(worker, r);
        worker(0);
        for (auto& t : th) t.join();
        for (int r = 0; r < n_devices; ++r)
            if (codes[r] != SCRC_OK) throw Error(codes[r], "rank " + std::to_string(r) + ": " + msgs[r]);
        *out = results[0].release();
    });
}

int scrc_grid_result_info(const scrc_grid_result* r, scrc_grid_info* info) {
    if (!r || !info) return SCRC_ERR_ARG;
    *info = r->info;
    return SCRC_OK;
}
int scrc_grid_result_penalty(const scrc_grid_result* r, int i, scrc_grid_penalty_info* info) {
    if (!r || !info || i < 0 || i >= (int)r->pens.size()) return SCRC_ERR_ARG;
    const PenaltyResult& p = r->pens[i];
    info->penalization = p.penalization; info->converged = p.converged ? 1 : 0; info->n_cycles_run = p.n_cycles_run;
    info->final_cycle_length = p.final_cycle_length; info->n_history = (int)p.k_history.size();
    info->n_records = (int)p.records.size();
    return SCRC_OK;
}
int scrc_grid_result_history(const scrc_grid_result* r, int i, int* k_history) {
    if (!r || !k_history || i < 0 || i >= (int)r->pens.size()) return SCRC_ERR_ARG;
    const PenaltyResult& p = r->pens[i];
    const size_t T = (size_t)r->info.T;
    for (size_t h = 0; h < p.k_history.size(); ++h) std::memcpy(k_history + h * T, p.k_history[h].data(), T * 4);
    return SCRC_OK;
}
int scrc_grid_result_record(const scrc_grid_result* r, int i, int rec, int* admm_iterations, int* n_selected,
                            unsigned char* models, int* votes, int* nnls_iterations) {
    if (!r || i < 0 || i >= (int)r->pens.size()) return SCRC_ERR_ARG;
    const PenaltyResult& p = r->pens[i];
    if (rec < 0 || rec >= (int)p.records.size()) return SCRC_ERR_ARG;
    const CycleRecord& c = p.records[rec];
    if (admm_iterations) std::memcpy(admm_iterations, c.admm_iterations.data(), c.admm_iterations.size() * 4);
    if (n_selected) std::memcpy(n_selected, c.n_selected.data(), c.n_selected.size() * 4);
    if (models) std::memcpy(models, c.models.data(), c.models.size());
    if (votes) { if (c.votes.empty()) return SCRC_ERR_ARG; std::memcpy(votes, c.votes.data(), c.votes.size() * 4); }
    if (nnls_iterations) {
        if (c.nnls_iterations.empty()) return SCRC_ERR_ARG;
        std::memcpy(nnls_iterations, c.nnls_iterations.data(), c.nnls_iterations.size() * 4);
    }
    return SCRC_OK;
}
int scrc_grid_result_records(const scrc_grid_result* r, int i, int* admm_iterations, int* n_selected, unsigned char* models) {
    if (!r || i < 0 || i >= (int)r->pens.size()) return SCRC_ERR_ARG;
    const PenaltyResult& p = r->pens[i];
    const size_t K = (size_t)r->info.K, KP = K * (size_t)r->info.P;
    for (size_t rec = 0; rec < p.records.size(); ++rec) {
        const CycleRecord& c = p.records[rec];
        if (admm_iterations) std::memcpy(admm_iterations + rec * K, c.admm_iterations.data(), K * 4);
        if (n_selected) std::memcpy(n_selected + rec * K, c.n_selected.data(), K * 4);
        if (models) std::memcpy(models + rec * KP, c.models.data(), KP);
    }
    return SCRC_OK;
}
static const FinalConfig* final_of(const scrc_grid_result* r, int i, int m) {
    if (!r || i < 0 || i >= (int)r->pens.size()) return nullptr;
    const PenaltyResult& p = r->pens[i];
    if (m < 0 || m >= (int)p.finals.size()) return nullptr;
    return &p.finals[m];
}
int scrc_grid_result_final(const scrc_grid_result* r, int i, int m, scrc_grid_final* out) {
    const FinalConfig* f = final_of(r, i, m);
    if (!f || !out) return SCRC_ERR_ARG;
    auto cp = [](void* dst, const void* src, size_t bytes) { if (dst && bytes) std::memcpy(dst, src, bytes); };
    cp(out->k_final, f->k_final.data(), f->k_final.size() * 4);
    cp(out->models, f->models.data(), f->models.size());
    cp(out->weights, f->weights.data(), f->weights.size() * 8);
    cp(out->signs, f->signs.data(), f->signs.size() * 8);
    cp(out->r2_test, f->r2_test.data(), f->r2_test.size() * 8);
    cp(out->resid_var, f->resid_var.data(), f->resid_var.size() * 8);
    cp(out->best_r2, f->best_r2.data(), f->best_r2.size() * 8);
    cp(out->best_r2_idx, f->best_r2_idx.data(), f->best_r2_idx.size() * 4);
    cp(out->n_selected, f->n_selected.data(), f->n_selected.size() * 4);
    if (out->r2_module) { if (f->r2_module.empty()) return SCRC_ERR_ARG; cp(out->r2_module, f->r2_module.data(), f->r2_module.size() * 8); }
    if (out->importance) { if (f->importance.empty()) return SCRC_ERR_ARG; cp(out->importance, f->importance.data(), f->importance.size() * 8); }
    return SCRC_OK;
}
int scrc_grid_result_coeffs(const scrc_grid_result* r, int i, int m, double* coeffs_out, int* sel_out) {
    const FinalConfig* f = final_of(r, i, m);
    if (!f || !f->have_coeffs) { scrc_set_last_error_("scrc_grid_result_coeffs: no coefficients kept for this configuration"); return SCRC_ERR_ARG; }
    return grid_guarded([&] {
        if (sel_out && !f->sel.empty()) std::memcpy(sel_out, f->sel.data(), f->sel.size() * 4);
        if (!coeffs_out || f->coeffs_n == 0) return;
        if (!f->coeffs_host.empty()) { std::memcpy(coeffs_out, f->coeffs_host.data(), (size_t)f->coeffs_n * 8); return; }
        SCRC_CUDA(cudaSetDevice(r->device));
        SCRC_CUDA(cudaMemcpy(coeffs_out, f->coeffs_dev.p, (size_t)f->coeffs_n * 8, cudaMemcpyDeviceToHost));
    });
}
int64_t scrc_grid_result_r2_removed_size(const scrc_grid_result* r, int i, int m) {
    const FinalConfig* f = final_of(r, i, m);
    return f ? (int64_t)f->r2_removed.size() : -1;
}
int scrc_grid_result_r2_removed(const scrc_grid_result* r, int i, int m, double* out) {
    const FinalConfig* f = final_of(r, i, m);
    if (!f || !out) return SCRC_ERR_ARG;
    if (!f->r2_removed.empty()) std::memcpy(out, f->r2_removed.data(), f->r2_removed.size() * 8);
    return SCRC_OK;
}
void scrc_grid_result_free(scrc_grid_result* r) {
    if (!r) return;
    if (r->device >= 0) cudaSetDevice(r->device);
    delete r;
}

}  // extern "C"
