/* b2s_host.h -- C entry points of the C++ host driver (case interface kept drop-in).
 *
 * The driver reads the GenCase output the reference solver reads (Case.xml + Case.bi4), owns a b2s_ctx
 * (include/b2s.h) and runs the JSphGpuSingle-style loop (src/source/JSphGpuSingle.cpp:950-1041), writing
 * dirout/data/Part_XXXX.bi4 + Part_Head.ibi4 like JSph::SavePartData (src/source/JSph.cpp:3062-3146).
 * `options` is a blank-separated string of the reference's own command-line switches
 * (src/source/JSphCfgRun.cpp:312-524), e.g. "-symplectic -wendland -viscoart:0.1 -ddt:2 -tmax:0.5";
 * unsupported switches are an error, never ignored.
 */
#ifndef B2S_HOST_H_
#define B2S_HOST_H_
#include "b2s.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct b2s_sim b2s_sim;
/* casepath: path of the case without extension (".xml" accepted). dirout may be NULL/"" (no PART output). */
int b2s_sim_load(const char* casepath, const char* dirout, const char* options, int device,
                 uint32_t extra_capacity, b2s_sim** out);
/* One rank of a slab-decomposed multi-GPU run (no reference counterpart: DualSPHysics v5.4 is single-GPU).  Every rank
 * reads the same case, derives the same balanced cut planes along x from the particle histogram, uploads ONLY its slab
 * and calls b2s_slab_config.  With mDBC the halo is two kernel radii wide (boundary ghosts recompute their ghost-node density locally). */
int b2s_sim_load_slab(const char* casepath, const char* options, int device, int rank, int nranks, b2s_sim** out);
/* Connects a slab rank to the others (csrc/slab_comm.cu): NCCL communicator from the 128-byte id rank 0 obtained with
 * b2s_slab_comm_id, then the first halo exchange.  Afterwards b2s_sim_run steps collectively (b2s_slab_step) and writes pieces. */
int b2s_sim_slab_connect(b2s_sim* sim, const uint8_t* nccl_id128);
/* PART output of a slab rank: every rank writes its own piece of each PART the way the reference's format provides for
 * (src/source/JPartDataBi4.cpp:99-127: Part_pRR_XXXX.bi4, PartInfo_pRR.ibi4, head values Piece / Npiece), rank 0 also Part_Head.ibi4
 * and PartFloatInfo.ibi4; no gather.  Call once after b2s_sim_load_slab, then b2s_sim_save_part on every rank at the same step. */
int b2s_sim_set_dirout(b2s_sim* sim, const char* dirout);
/* cuts[nranks+1] (cell planes along x) and the particle count of the densest plane (sizes the exchange buffers). */
int b2s_sim_slab_info(b2s_sim* sim, uint32_t* cuts, uint32_t ncuts, uint32_t* max_plane_particles);
int b2s_sim_free(b2s_sim* sim);
int b2s_sim_ctx(b2s_sim* sim, b2s_ctx** out);            /* borrowed: freed by b2s_sim_free */
int b2s_sim_params(b2s_sim* sim, b2s_params* out);
/* Number of the next PART and the excluded particles written to PartOut so far: [0] all, [1] position, [2] density, [3] velocity
 * (what JSph::ShowResume / GetResInfo report, src/source/JSph.cpp:3471-3494,3522-3560). */
int b2s_sim_stats(b2s_sim* sim, uint32_t* parts, uint32_t* excluded4);
int b2s_sim_save_part(b2s_sim* sim);                     /* writes the next Part_XXXX.bi4 */
/* Runs until TimeMax or nsteps_break steps (0: use -nsteps option / none). */
int b2s_sim_run(b2s_sim* sim, uint32_t nsteps_break, double* time_out, uint32_t* steps_out);
/* PartMotionRef.ibi4 (src/source/JDsPartMotionSave.cpp:118-235, JComputeMotionRef.cpp:127-236, JPartMotRefBi4Save.cpp:149-292) on host
 * arrays: picks the three reference particles of each body (moving ones first) from the first set of positions and writes the head
 * plus one PART item per set into dirdata.  idp[n] ascending, pos[nparts][n][3].  The driver calls the same code at every PART;
 * this entry point lets the file be checked against the reference's own without a GPU. */
int b2s_host_motionref_write(const char* dirdata, uint32_t nbodies, const uint8_t* floating, const uint16_t* mkbound,
                             const uint32_t* begin, const uint32_t* count, double timeout, uint32_t mkboundfirst,
                             uint32_t nparts, uint32_t n, const uint32_t* idp, const double* pos,
                             const double* timestep, const uint32_t* step);
/* The driver's XML reader (csrc/host/xml_lite.cpp, in place of the reference's tinyxml/JXml) on any file: number of elements and
 * of attributes of the whole tree.  Test entry point: the reader is checked against every XML file the reference ships. */
int b2s_host_xml_stats(const char* file, uint32_t* elements, uint32_t* attributes);
const char* b2s_host_last_error(void);
#ifdef __cplusplus
}
#endif
#endif
