// Field pressure of many sources at 10^5..10^6 observation points (C ABI: cnld_b200_pfield_*).
//
// Replaces the loop of array_patch_pres_imp_resp (cnld/bem_pressure.pyx:103-112) -- for every
// frequency and every source patch j: sfr[j] = mesh_pres_vector(r, k) . disp[j] -- for ALL field
// points at once.  Per wavenumber, per super-chunk of observation points (one wave of k_pvec_tiles
// CTAs): PV = pressure vectors in node space (exp(-jkr)/r evaluated once per point pair), then
// p[src][obs] = PV . disp[src] on the DMMA ZGEMM; the chunk's results start their D2H copy on a
// second stream while the next chunk is computed.
#include "ctx.cuh"

#include <algorithm>

struct cnld_pfield {
  cnld_ctx *ctx = nullptr;
  uint nobs = 0, nsrc_max = 0, mc = 0;   // mc = observation points per super-chunk
  size_t lda = 0;
  double *d_obs = nullptr;
  cplx *d_PV = nullptr, *d_disp = nullptr, *d_out = nullptr;
  cudaStream_t st = nullptr, st_copy = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // begin, end, chunk done, copies done
  std::vector<cudaEvent_t> ev_k;                              // per chunk: pvec begin / pvec end / gemm end
  void *reg_ptr = nullptr;
  double times[4] = {0, 0, 0, 0};
};

namespace {
__global__ void k_pf_gather(uint nsrc, uint nidx, const uint *__restrict__ idx, const cplx *__restrict__ src, size_t ld,
                            cplx *dst) {
  const uint i = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (i < nidx && s < nsrc) dst[(size_t)s * nidx + i] = src[(size_t)s * ld + idx[i]];
}

int ensure_plan(cnld_ctx *c) {
  if (c->pplan_ready && c->pplan.variant != g_pvec_variant) { pres_plan_free(c->pplan); c->pplan_ready = false; }
  if (c->pplan_ready) return 0;
  if (pres_plan_build(c->bem, c->h_x.data(), c->h_t.data(), c->pplan)) return -1;
  c->pplan_ready = true;
  return 0;
}

int pf_run(cnld_pfield *pf, double k, double cs, double rho, const cnld_field *disp, uint nsrc, cnld_field *p) {
  cnld_ctx *c = pf->ctx;
  CNLD_CK(cudaSetDevice(c->bem.device));
  if (nsrc == 0 || nsrc > pf->nsrc_max) { set_error("pfield: nsrc %u outside 1..%u", nsrc, pf->nsrc_max); return -1; }
  const size_t nv = c->bem.nv, nobs = pf->nobs;
  CNLD_CK(cudaEventRecord(pf->ev[0], pf->st));
  if (disp)
    CNLD_CK(cudaMemcpyAsync(pf->d_disp, disp, sizeof(cplx) * nsrc * nv, cudaMemcpyHostToDevice, pf->st));
  const uint nchunks = (pf->nobs + pf->mc - 1) / pf->mc;
  while (pf->ev_k.size() < 3 * (size_t)nchunks) {
    cudaEvent_t e;
    CNLD_CK(cudaEventCreate(&e));
    pf->ev_k.push_back(e);
  }
  int rc = 0;
  for (uint ch = 0; ch < nchunks && !rc; ch++) {
    const uint o0 = ch * pf->mc, mc = std::min(pf->mc, pf->nobs - o0);
    // the previous chunk's D2H reads d_out only, the PV buffer is free as soon as its ZGEMM is done
    // (same stream), so there is nothing to wait for here
    CNLD_CK(cudaEventRecord(pf->ev_k[3 * ch], pf->st));
    if (pres_pvec_tiles(pf->st, c->pplan, c->bem, pf->d_obs + 3 * (size_t)o0, mc, k, cs, rho, pf->d_PV, pf->lda)) { rc = -1; break; }
    CNLD_CK(cudaEventRecord(pf->ev_k[3 * ch + 1], pf->st));
    if (pres_contract(pf->st, c->bem, pf->d_PV, pf->lda, mc, pf->d_disp, nsrc, pf->d_out + o0, nobs)) { rc = -1; break; }
    CNLD_CK(cudaEventRecord(pf->ev_k[3 * ch + 2], pf->st));
    if (p) {
      CNLD_CK(cudaStreamWaitEvent(pf->st_copy, pf->ev_k[3 * ch + 2], 0));
      CNLD_CK(cudaMemcpy2DAsync(reinterpret_cast<cplx *>(p) + o0, sizeof(cplx) * nobs, pf->d_out + o0, sizeof(cplx) * nobs,
                                sizeof(cplx) * mc, nsrc, cudaMemcpyDeviceToHost, pf->st_copy));
    }
  }
  if (!rc && p) {
    CNLD_CK(cudaEventRecord(pf->ev[3], pf->st_copy));
    CNLD_CK(cudaStreamWaitEvent(pf->st, pf->ev[3], 0));
  }
  CNLD_CK(cudaEventRecord(pf->ev[1], pf->st));
  cudaError_t e = cudaStreamSynchronize(pf->st);
  cudaError_t e2 = cudaStreamSynchronize(pf->st_copy);
  if (e != cudaSuccess || e2 != cudaSuccess) {
    set_error("pfield: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return -1;
  }
  if (rc) return rc;
  float ms = 0;
  pf->times[0] = pf->times[1] = 0;
  for (uint ch = 0; ch < nchunks; ch++) {
    cudaEventElapsedTime(&ms, pf->ev_k[3 * ch], pf->ev_k[3 * ch + 1]); pf->times[0] += ms * 1e-3;
    cudaEventElapsedTime(&ms, pf->ev_k[3 * ch + 1], pf->ev_k[3 * ch + 2]); pf->times[1] += ms * 1e-3;
  }
  cudaEventElapsedTime(&ms, pf->ev[0], pf->ev[1]);
  pf->times[2] = ms * 1e-3;
  pf->times[3] = (double)nchunks;
  return 0;
}
}  // namespace

extern "C" {

cnld_pfield *cnld_b200_pfield_create(cnld_ctx *c, const double *obs, cnld_uint nobs, cnld_uint nsrc_max) {
  if (!c || !obs || !nobs || !nsrc_max) { set_error("pfield_create: bad arguments"); return nullptr; }
  if (cudaSetDevice(c->bem.device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
  if (ensure_plan(c)) return nullptr;
  cnld_pfield *pf = new cnld_pfield();
  pf->ctx = c; pf->nobs = nobs; pf->nsrc_max = nsrc_max;
  const size_t nv = c->bem.nv;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->bem.device);
  const uint tile = pres_tile_obs();
  uint wave = (uint)nsm * pres_tile_ctas() * tile;           // one full wave of k_pvec_tiles CTAs
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const size_t fixed = sizeof(cplx) * ((size_t)nsrc_max * nobs + (size_t)nsrc_max * nv) + 24 * (size_t)nobs;
  const size_t budget = free_b > fixed + (1ull << 30) ? (free_b - fixed - (1ull << 30)) / 2 : 0;
  while (wave > tile && sizeof(cplx) * (size_t)wave * nv > budget) wave -= tile;
  pf->mc = std::min(wave, ((nobs + tile - 1) / tile) * tile);
  pf->lda = pf->mc;
  bool ok = cudaMalloc(&pf->d_obs, sizeof(double) * 3 * (size_t)nobs) == cudaSuccess &&
            cudaMalloc(&pf->d_PV, sizeof(cplx) * pf->lda * nv) == cudaSuccess &&
            cudaMalloc(&pf->d_disp, sizeof(cplx) * (size_t)nsrc_max * nv) == cudaSuccess &&
            cudaMalloc(&pf->d_out, sizeof(cplx) * (size_t)nsrc_max * nobs) == cudaSuccess &&
            cudaMemcpy(pf->d_obs, obs, sizeof(double) * 3 * (size_t)nobs, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemset(pf->d_disp, 0, sizeof(cplx) * (size_t)nsrc_max * nv) == cudaSuccess &&
            cudaStreamCreateWithFlags(&pf->st, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&pf->st_copy, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 4 && ok; i++) ok = cudaEventCreate(&pf->ev[i]) == cudaSuccess;
  if (!ok) {
    set_error("pfield_create: %s (nobs %u, nsrc %u, chunk %u)", cudaGetErrorString(cudaGetLastError()), nobs, nsrc_max, pf->mc);
    cnld_b200_pfield_destroy(pf);
    return nullptr;
  }
  return pf;
}

void cnld_b200_pfield_destroy(cnld_pfield *pf) {
  if (!pf) return;
  cudaSetDevice(pf->ctx->bem.device);
  if (pf->st) cudaStreamSynchronize(pf->st);
  if (pf->st_copy) cudaStreamSynchronize(pf->st_copy);
  if (pf->reg_ptr) cudaHostUnregister(pf->reg_ptr);
  cudaFree(pf->d_obs); cudaFree(pf->d_PV); cudaFree(pf->d_disp); cudaFree(pf->d_out);
  for (auto e : pf->ev) if (e) cudaEventDestroy(e);
  for (auto e : pf->ev_k) cudaEventDestroy(e);
  if (pf->st) cudaStreamDestroy(pf->st);
  if (pf->st_copy) cudaStreamDestroy(pf->st_copy);
  delete pf;
}

int cnld_b200_pfield_pin(cnld_pfield *pf, void *ptr, size_t bytes) {
  if (!pf) { set_error("null pfield"); return -1; }
  CNLD_CK(cudaSetDevice(pf->ctx->bem.device));
  if (pf->reg_ptr) { cudaHostUnregister(pf->reg_ptr); pf->reg_ptr = nullptr; }
  if (!ptr || !bytes) return 0;
  if (cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaHostRegister failed for %zu bytes", bytes);
    return -1;
  }
  pf->reg_ptr = ptr;
  return 0;
}

int cnld_b200_pfield_run(cnld_pfield *pf, double k, double c, double rho, const cnld_field *disp, cnld_uint nsrc,
                         cnld_field *p) {
  if (!pf || !disp) { set_error("pfield_run: null argument"); return -1; }
  return pf_run(pf, k, c, rho, disp, nsrc, p);
}

int cnld_b200_pfield_run_device(cnld_pfield *pf, double k, double c, double rho, cnld_uint nsrc) {
  if (!pf) { set_error("null pfield"); return -1; }
  return pf_run(pf, k, c, rho, nullptr, nsrc, nullptr);
}

int cnld_b200_pfield_fetch(cnld_pfield *pf, cnld_uint nsrc, const cnld_uint *idx, cnld_uint nidx, cnld_field *p) {
  if (!pf || !p || nsrc > pf->nsrc_max) { set_error("pfield_fetch: bad arguments"); return -1; }
  CNLD_CK(cudaSetDevice(pf->ctx->bem.device));
  if (!idx) {
    CNLD_CK(cudaMemcpy(p, pf->d_out, sizeof(cplx) * (size_t)nsrc * pf->nobs, cudaMemcpyDeviceToHost));
    return 0;
  }
  for (uint i = 0; i < nidx; i++)
    if (idx[i] >= pf->nobs) { set_error("pfield_fetch: index out of range"); return -1; }
  if (!nidx || !nsrc) return 0;
  uint *d_idx = nullptr;
  cplx *d_g = nullptr;
  int rc = -1;
  do {
    if (cudaMalloc(&d_idx, sizeof(uint) * nidx) != cudaSuccess || cudaMalloc(&d_g, sizeof(cplx) * (size_t)nsrc * nidx) != cudaSuccess) break;
    if (cudaMemcpyAsync(d_idx, idx, sizeof(uint) * nidx, cudaMemcpyHostToDevice, pf->st) != cudaSuccess) break;
    k_pf_gather<<<dim3((nidx + 127) / 128, nsrc), 128, 0, pf->st>>>(nsrc, nidx, d_idx, pf->d_out, pf->nobs, d_g);
    count_launch();
    if (cudaMemcpyAsync(p, d_g, sizeof(cplx) * (size_t)nsrc * nidx, cudaMemcpyDeviceToHost, pf->st) != cudaSuccess) break;
    if (cudaStreamSynchronize(pf->st) != cudaSuccess) break;
    rc = 0;
  } while (0);
  if (rc) set_error("pfield_fetch: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_idx); cudaFree(d_g);
  return rc;
}

int cnld_b200_pres_plan_stats(cnld_uint nv, cnld_uint nt, const double *x, const cnld_uint *t, cnld_uint *counts,
                              cnld_uint *order) {
  if (!x || !t || !counts) { set_error("null argument"); return -1; }
  for (size_t i = 0; i < 3 * (size_t)nt; i++)
    if (t[i] >= nv) { set_error("triangle vertex index out of range"); return -1; }
  return pres_plan_stats(nv, nt, x, t, counts, order);
}

int cnld_b200_set_pvec_variant(int v) { return pres_set_variant(v); }

int cnld_b200_pfield_info(cnld_pfield *pf, double *out) {
  if (!pf || !out) { set_error("null argument"); return -1; }
  out[0] = pf->times[0]; out[1] = pf->times[1]; out[2] = pf->times[2]; out[3] = pf->times[3];
  out[4] = pf->mc; out[5] = pf->ctx->pplan.nchunk; out[6] = pf->ctx->pplan.nentries;
  return 0;
}

}  // extern "C"
