// sgdx_handle.h -- the handle behind include/sgdx.h, shared by the C-ABI translation units
// (sgdx_api.cu: assignment path; sgdx_post.cu: the stages either side of it).
#pragma once
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sgdx.h"
#include "sgdx_kernels.h"

namespace sgdx {

extern thread_local std::string g_err;
int fail(int code, const char *fmt, ...);

#define CU(call)                                                                                                \
    do {                                                                                                        \
        cudaError_t e__ = (call);                                                                               \
        if (e__ != cudaSuccess)                                                                                 \
            return ::sgdx::fail(SGDX_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,  \
                                __LINE__);                                                                      \
    } while (0)

struct Slot {
    cudaStream_t own = nullptr;
    cudaStream_t stream = nullptr;  // own or caller-provided
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    uint8_t *d_in = nullptr;
    uint16_t *d_idx = nullptr;
    uint8_t *d_dist = nullptr;
    uint64_t cap = 0;  // reads
    bool timed = false;
    unsigned long long *h_unm_stats = nullptr;  // pinned copy of UnmatchedState::d_stats after the slot's last batch
    bool unm_pending = false;                   // ... and whether one is in flight
    uint8_t *d_extract = nullptr;               // scratch of sgdx_match_segments_device
    uint64_t extract_cap = 0;
    void *d_route = nullptr;                    // scratch of sgdx_route_device (per-warp histogram matrix)
    uint64_t route_cap = 0;                     // bytes
    uint32_t *d_qmasks = nullptr;               // shifted segment masks of the last read structure this slot
    uint64_t qmasks_cap = 0;                    //   counted qualities for (words), and that structure
    std::vector<uint32_t> qmasks_key;
    uint4 *d_cexp = nullptr;                    // scratch of the compact-input general path (16-byte records)
    uint64_t cexp_cap = 0;
    uint8_t *d_dense = nullptr;                 // dense records of a batch as they arrive (sgdx_match_dense_batch)
    uint64_t dense_cap = 0;                     // bytes
    uint32_t *d_exc_idx = nullptr;              // ... and its exception list
    void *d_exc_rec = nullptr;                  // (8- or 16-byte records: SGDX_DENSE_EXC_BYTES)
    uint64_t exc_cap = 0;                       // entries
};

struct UnmSlot;

// most-frequent-unmatched collection (sgdx_post.cu); the table lives on the device
struct UnmatchedState {
    bool on = false;
    uint64_t max_map_size = 0, downsize_to = 0;
    uint64_t cap_slots = 0;   // power of two
    uint64_t key_limit = 0;   // new keys are refused (and counted as dropped) beyond this many
    UnmSlot *d_slots = nullptr;             // 32-byte slots: key + count (sgdx_post.cu)
    unsigned long long *d_stats = nullptr;  // kUnmStats words, see sgdx_post.cu
    uint8_t *d_odd = nullptr;               // barcodes with a byte outside ACGTN, 32 bytes each
    uint64_t odd_cap = 0;
    std::map<std::string, uint64_t> odd;    // host-side counts of those
    ulonglong2 *d_sel_keys = nullptr;       // selection scratch
    unsigned long long *d_sel_counts = nullptr;
    uint64_t sel_cap = 0;
    unsigned long long *d_hist = nullptr;
    uint64_t downsizes = 0;
    int ctas_per_sm = 1;                    // resident CTAs of sgdx_unmatched_count per SM
};

struct QualState {
    unsigned long long *d_counters = nullptr;  // [(S+1)][5] + 1 word of short reads
};

}  // namespace sgdx

struct sgdx_handle {
    sgdx_config cfg;  // effective (AUTO resolved)
    uint32_t mode = sgdx::kModeHamming;
    int sms = 148;
    sgdx::DevParams base{};  // everything but the per-batch pointers
    uint2 *d_table = nullptr;
    uint16_t *d_seed_hdr = nullptr;
    uint16_t *d_seed_ids = nullptr;
    uint32_t *d_ex_slots = nullptr, *d_ex_ascii = nullptr, *d_nb_slots = nullptr;
    bool seeded = false;  // launches use sgdx_match_seeded
    uint64_t *d_ph_hop_slots = nullptr;
    uint16_t *d_ph_hop_disp = nullptr;
    uint16_t *d_ph_ms_hdr = nullptr, *d_ph_ms_ids = nullptr;  // short seed index of the perfect-hash kernel (PhParams::ms_*)
    uint8_t *d_ph_nn = nullptr;                               // ... and PhParams::nn
    uint16_t *d_ph_disp = nullptr, *d_ph_slots = nullptr, *d_ph_b_disp = nullptr, *d_ph_b_slots = nullptr;  // perfect-hash kernel (sgdx_ph.cu); ph.slots == nullptr: not available
    sgdx::PhParams ph{};
    uint32_t ph_keys = 0;
    bool single = false;  // compact records run sgdx_match_single
    uint4 *d_table_long = nullptr;  // barcode_len > 32: {loA, hiA, loB, hiB} per sample, the handle runs sgdx_match_long
    sgdx::HopSlot *d_hash1 = nullptr, *d_hash2 = nullptr;
    uint16_t *d_hop_lut1 = nullptr, *d_hop_lut2 = nullptr;
    unsigned long long *d_hop_fwd = nullptr, *d_hop_rev = nullptr;
    ulonglong2 *d_hop_list = nullptr;            // scratch of the hop export: the non-zero cells, compacted on the device
    unsigned long long *d_hop_stat = nullptr;
    uint64_t hop_list_cap = 0;
    std::mutex hop_mu;
    unsigned long long *d_counters = nullptr;
    std::vector<std::string> keys1, keys2;  // distinct index1 / index2 halves (ASCII), id order
    std::vector<sgdx::Slot> slots;
    std::atomic<uint64_t> launches{0};
    // stages either side of the assignment (sgdx_post.cu)
    std::mutex post_mu;  // orders batch launches against a downsize of the unmatched table
    sgdx::UnmatchedState unm;
    sgdx::QualState qual;
};

namespace sgdx {
// hooks of sgdx_post.cu called by sgdx_api.cu
int post_after_match(sgdx_handle *h, Slot &s, const uint8_t *d_ascii, const uint64_t *d_compact, const uint4 *d_packed,
                     const uint16_t *d_idx, uint64_t n);
int post_on_sync(sgdx_handle *h, Slot &s);
int post_reset(sgdx_handle *h);
void post_destroy(sgdx_handle *h);
}  // namespace sgdx
