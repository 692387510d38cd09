// Packed weight layout shared by the host C-ABI and the kernels ("Data layout").
//
// The reference keeps 2*nq separate nn.Linear modules (rqae/model.py:29-36).  The kernels
// stream them as nq+1 *stages*; stage s holds exactly what one pass of the fused loop needs:
//     W_out[s-1] rows + b_out[s-1]   (out-projection of the code chosen at layer s-1)
//     W_in[s] columns                (in-projection of layer s)
// with zeros for W_out[-1], b_out[-1] and W_in[nq].  A stage is cut into CH chunks, one
// bulk-TMA copy each.  A CTA has 256 compute threads; thread t owns elements
// d = j*256 + t (j < E) of the D axis, so inside a chunk (JC = E/CH values of j):
//     [ float4 w_out4 [JC][256] | float4 w_in4 [JC][256] | float b_out [JC][256] ]
// and a warp's 128-bit shared loads are 512 contiguous bytes (conflict free).
#pragma once
#include <stddef.h>
#include <stdint.h>

#define RQ_GROUP_THREADS 256 /* compute threads per CTA */
#define RQ_BYTES_PER_ELEM 36 /* float4 + float4 + float */
#define RQ_HDR_BYTES 256
#define RQ_SMEM_ROWS 640 /* rows of the de-duplicated search table kept in shared memory */
#define RQ_CAN_MAX 16     /* canonical rows (c0 >= c1 >= c2 >= c3 >= 0) of a sign/permutation-symmetric table */
#define RQ_NPERM 24
#define RQ_NSIGN 16

/* ring geometry of the Gemma-2-2B shape: chunks per stage and ring slots (tuning knobs) */
#ifndef RQ_E9_CH
#define RQ_E9_CH 1
#define RQ_E9_NSLOT 2
#endif
#define RQ_E9_DEC_NSLOT (RQ_E9_CH == 1 ? 4 : 8)

struct RqShape {
  int E;     /* elements per thread: D_pad = 256 * E */
  int EC;    /* register block (elements per inner block) */
  int CH;    /* chunks per stage */
  int NSLOT; /* ring slots */
  int TG;    /* tokens per phase (a CTA iteration handles 2 phases = 2*TG tokens) */
};

/* Supported instantiations, smallest first. Returns 0 on success. */
static inline int rq_pick_shape(int D, struct RqShape* s) {
  static const struct RqShape table[] = {
      {1, 1, 1, 4, 8},    /* D <=  256 */
      {3, 3, 1, 4, 8},    /* D <=  768 */
      {6, 3, 2, 6, 8},    /* D <= 1536 */
      {9, 3, RQ_E9_CH, RQ_E9_NSLOT, 8}, /* D <= 2304  (Gemma-2-2B) */
      {14, 2, 7, 10, 6},  /* D <= 3584  (Gemma-2-9B) */
  };
  for (size_t i = 0; i < sizeof(table) / sizeof(table[0]); i++) {
    if (D <= table[i].E * RQ_GROUP_THREADS) {
      *s = table[i];
      return 0;
    }
  }
  return 1;
}

/* D-split cluster variant of the forward kernel (rq_forward.cuh, CS > 1): a cluster of CS CTAs works on ONE unit of
 * 16 tokens, CTA c owning elements d = (c * E/CS + j) * 256 + t.  Its weights are a second copy of the stages cut into
 * CS chunks, chunk c = CTA c's slice.  0 = the shape has no cluster variant. */
static inline int rq_cluster_size(int E) { return E == 14 ? 2 : 0; }

struct RqLayout {
  size_t off_bin;    /* float4[nq+1]          in-projection biases                    */
  size_t off_cbt;    /* float4[KT]            de-duplicated search table (shared mode) */
  size_t off_map;    /* uint16[KT]            search row -> lowest original index      */
  size_t off_tp;     /* float4[24][RQ_CAN_MAX] canonical rows, coordinates arranged per magnitude order */
  size_t off_map3;   /* uint16[16][24][RQ_CAN_MAX] (signs, order, canonical row) -> lowest original index */
  size_t off_stage;  /* (nq+1) stages                                                  */
  size_t off_stage_cl; /* (nq+1) stages in the cluster variant's chunking (0 if none)  */
  size_t stage_bytes;
  size_t chunk_bytes;
  size_t total;
  int KT; /* K rounded up to a multiple of 32 */
};

static inline size_t rq_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static inline void rq_layout(int nq, int K, const struct RqShape* s, struct RqLayout* L) {
  L->KT = (K + 31) / 32 * 32;
  L->off_bin = RQ_HDR_BYTES;
  L->off_cbt = rq_align_up(L->off_bin + (size_t)(nq + 1) * 16, 256);
  L->off_map = rq_align_up(L->off_cbt + (size_t)L->KT * 16, 256);
  L->off_tp = rq_align_up(L->off_map + (size_t)L->KT * 2, 256);
  L->off_map3 = L->off_tp + (size_t)RQ_NPERM * RQ_CAN_MAX * 16;
  L->off_stage = rq_align_up(L->off_map3 + (size_t)RQ_NSIGN * RQ_NPERM * RQ_CAN_MAX * 2, 1024);
  L->stage_bytes = (size_t)s->E * RQ_GROUP_THREADS * RQ_BYTES_PER_ELEM;
  L->chunk_bytes = L->stage_bytes / s->CH;
  L->total = L->off_stage + (size_t)(nq + 1) * L->stage_bytes;
  L->off_stage_cl = 0;
  if (rq_cluster_size(s->E)) {
    L->off_stage_cl = rq_align_up(L->total, 1024);
    L->total = L->off_stage_cl + (size_t)(nq + 1) * L->stage_bytes;
  }
}

/* Device-resident header at offset 0 of the packed buffer. */
struct RqHeader {
  int kd_pad;    /* rows in the search table (multiple of 32) */
  int kd;        /* distinct rows */
  int can_rows;   /* canonical rows, padded to a multiple of 4; 0 = table not symmetric, canonical search unused */
  float thr_tiny; /* canonical search needs every |z_i| >= thr_tiny * |z|                     */
  float thr_gap;  /*   ... the best score z.c leading the runner-up by more than thr_gap * |z| */
  float thr_sep;  /*   ... and any two |z_i| differing by at least thr_sep * |z|              */
  int reserved[58];
};
