// Internal interface of the tensor-memory DCN kernel (dcn2.cu); reached through b200_igemm_create / _launch with
// mode B200_IGEMM_DCN_TMEM.
#pragma once
#include "common.cuh"
#include "internal.h"

namespace b200 {

struct Dcn2Params {
    int batch, in_h, in_w, in_c, out_h, out_w, out_c, stride, flags;
    int block_n;  // == n_pad: 128 or 256 (the layer's whole output width is one UMMA N)
    int tile_w, tile_h, tiles_x, tiles_y, num_tiles;
    int k_blocks, cchunks;
    int x_ld;
    int b_stages, epi_bufs, nacc;
    int rec_lead;  // weight k-block of a tile at which the loader requests the NEXT tile's record box
    const __nv_bfloat16* x;
    const float* scale;
    const float* shift;
};

struct Dcn2Tmaps {
    CUtensorMap b;    // weights [n_pad][K], box {64, block_n}, 128B swizzle
    CUtensorMap out;  // output NHWC, box {64, tile_w, tile_h, 1}, 128B swizzle
    CUtensorMap rec;  // sampling records [N][H][W][36 x int32], box {36, tile_w, tile_h, 1}, no swizzle
};

struct Dcn2Plan {
    Dcn2Tmaps tm;
    Dcn2Params p;
    int grid, smem_bytes;
};

// `pick_tile_w` and the tensor-map encoder live in igemm.cu
int dcn2_supported(const b200_igemm_desc* d);
int dcn2_create(const b200_igemm_desc* d, Dcn2Plan** out);
int dcn2_launch(const Dcn2Plan* plan, void* stream);
void dcn2_destroy(Dcn2Plan* plan);

int encode_map_typed(CUtensorMap* m, int dtype /*0 = bf16, 1 = int32*/, int swizzle128, const void* base, int rank,
                     const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, const char* what);

}  // namespace b200
