/*
 * mkb_device_abi.h — structs shared by the host runtime (mkb_runtime.cu) and
 * the generated model kernels (myokit_b200/kernelgen.py). The runtime embeds
 * this text and hands it to NVRTC as the virtual header "mkb_device_abi.h", so
 * there is exactly one definition.
 *
 * Layout in HBM (one GPU / one row slab):
 *   state   n_state planes of `stride` Reals, structure-of-arrays:
 *           state[k * stride + cid], cid = ix + iy * nx (x fastest, as the
 *           reference's cell ids, myokit/_sim/openclsim.cl:316). `stride` is
 *           the cell count rounded up to 32 elements so every plane starts
 *           128-byte aligned.
 *   V       the membrane-potential plane is double-buffered (v_in = V(t),
 *           v_out = V(t+dt)); the fused kernel reads neighbours' V(t) while
 *           writing V(t+dt). The reference needs two kernels and an idiff
 *           round trip for the same effect (openclsim.c:1066-1096).
 *   idiff   one plane, written only on logged steps (flag bit 0)
 *   inter   n_inter planes of logged intermediary variables (same flag)
 *   field   n_field planes of per-cell constants (set_field)
 *   gx, gy  conductance fields in the reference's layout:
 *           gx[(ny, nx-1)], gy[(ny-1, nx)] row-major (openclsim.py:1380-1384)
 */
#ifndef MKB_DEVICE_ABI_H
#define MKB_DEVICE_ABI_H

#define MKB_FLAG_STORE_AUX 1u   /* write idiff + logged intermediaries */

/* One entry per time step in the device schedule ring. Host doubles; the
 * kernel casts to Real exactly like openclsim.c:1063,1148,1155. */
struct MkbStepParams {
    double time;
    double dt;
    double pace;
    unsigned int flags;
    unsigned int step;      /* 1-based index of this step within the run */
};

struct MkbGridArgs {
    void* state;                    /* Real[n_state][stride] */
    void* idiff;                    /* Real[stride] */
    void* inter;                    /* Real[n_inter][stride] */
    const void* field;              /* Real[n_field][stride] */
    const void* gx_field;           /* Real[ny][nx-1] or null */
    const void* gy_field;           /* Real[ny-1][nx] or null */
    const unsigned char* paced_mask;/* [n] or null (rectangle in use) */
    /* connections as CSR over cells (deterministic gather) */
    const unsigned long long* csr_row;  /* [n+1] */
    const unsigned int* csr_col;        /* [nnz] */
    const void* csr_g;                  /* Real[nnz] */
    /* Partitioned connection graphs (multi-GPU): CSR columns >= n are ghost
     * cells owned by other GPUs; their V(t_step) sits in ghost[(step % 3) *
     * n_ghost + (col - n)], pushed there by the owners (peer stores). */
    const void* ghost;                  /* Real[3][n_ghost] or null */
    unsigned long long n_ghost;
    /* Every block waits until ghost_flags[ghost_import[k]] >= step for all k
     * (one flag per exporting GPU, raised after its stores for `step`). */
    const unsigned int* ghost_flags;
    const unsigned int* ghost_import;   /* [n_ghost_import] */
    unsigned long long n_ghost_import;
    /* Row-slab sharding (multi-GPU), all null on a single GPU.
     * halo_lo / halo_hi: THIS GPU's ghost rows, Real[3][nx] each (slot =
     *   step % 3): V(t_step) of global row iy_offset-1 / iy_offset+ny,
     *   written by the neighbouring GPU's step kernel with peer stores over
     *   NVLink; flag_lo / flag_hi: unsigned[n column blocks], the last step
     *   whose row segment has landed (written by the neighbour after a
     *   system-scope fence).
     *   Why three slots: GPU r at step n needs its neighbour's V(t_n), which
     *   the neighbour writes during its step n-1 into slot n % 3. The
     *   neighbour cannot start step n+1 before r has delivered V(t_{n+1})
     *   (r's step n), so while r reads slot n % 3 the neighbour is writing
     *   at most slot (n+1) % 3 or — one step ahead — (n+2) % 3: never the
     *   slot being read.
     * peer_*: the same four arrays of the neighbouring GPUs, mapped into this
     *   process (cudaIpcOpenMemHandle or direct peer access): this GPU's
     *   first row goes to the lower neighbour's halo_hi, its last row to the
     *   upper neighbour's halo_lo. */
    const void* halo_lo;
    const void* halo_hi;
    const unsigned int* flag_lo;
    const unsigned int* flag_hi;
    void* peer_lo_halo_hi;
    void* peer_hi_halo_lo;
    unsigned int* peer_lo_flag_hi;
    unsigned int* peer_hi_flag_lo;
    unsigned int* halo_error;       /* set when a ghost-row wait times out */
    unsigned long long nx;          /* cells in x */
    unsigned long long ny;          /* rows in this slab */
    unsigned long long stride;      /* plane stride in elements */
    unsigned long long iy_offset;   /* first global row of this slab */
    unsigned long long ny_global;   /* rows in the whole grid */
    double gx, gy;                  /* homogeneous conductances */
    long long pace_x0, pace_x1;     /* paced rectangle [x0,x1) x [y0,y1) */
    long long pace_y0, pace_y1;
    /* Junction with a second grid stepped in lockstep (fibre-tissue,
     * myokit/_sim/openclsim.cl:601-628). Cells (jx, jy0 + k), k < jn, of this
     * grid are coupled with conductance jg to element joff + k * jstride of the
     * other grid's V(t) plane: junction_v0 when this grid reads its first V
     * plane, junction_v1 when it reads its second (both grids swap planes
     * every step). Kernels generated with junction='fiber' add
     * jg * (V - Vother) to idiff, junction='tissue' subtract jg * (Vother - V).
     * All zero / null on ordinary simulations. */
    const void* junction_v0;
    const void* junction_v1;
    double jg;
    unsigned long long jx, jy0, jn, joff, jstride;
    /* Kernels whose consecutive steps overlap (kernelgen overlap=True): one
     * counter per thread block of the launch grid, [row block][column block]:
     * the last step that block has completed (stores released). Zeroed when
     * a run starts counting its steps from 1. Null otherwise. */
    unsigned int* tile_done;
    /* Streaming kernels (kernelgen stream=True): TMA descriptors
     * (cuTensorMapEncodeTiled) of the two membrane-potential planes as 2-d
     * tensors [ny][nx] with the box the kernel was generated for: tmap[0] the
     * plane inside `state`, tmap[1] the second V plane. A kernel that uses them
     * declares its grid argument __grid_constant__. Zero elsewhere. */
    alignas(64) unsigned long long tmap[2][16];
    /* Staged kernels (kernelgen stage=True): TMA descriptor of the state
     * planes as a 3-d tensor [n_state][ny][nx] (strides: row nx, plane
     * `stride`) with a box of one thread-block tile of one plane. Loads fill
     * cells outside the grid with zeros, stores clip them. Zero elsewhere. */
    alignas(64) unsigned long long tmap_state[16];
};

#endif
