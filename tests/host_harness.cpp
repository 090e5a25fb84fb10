// CPU unit-test harness: compiles the product's per-macroblock entropy logic (entropy.cuh,
// h264_core.cuh) as ordinary C++ and drives it the way the CUDA kernels do (count pass, prefix sum,
// scatter pass; binarise then serial arithmetic coder; parallel-rule emulation prevention).
// TEST ONLY: it lets `pytest -m "not gpu"` compare this logic with the oracle without a GPU.
// It is not linked into the product library and does not link the oracle.
#include "../cedarx_h264_encoder_b200/csrc/entropy.cuh"

#include <cstring>
#include <vector>

using namespace cedar;

extern "C" {

// One slice (macroblock rows [row0, row1) of a picture whose slices are srows rows tall) through the
// count / prefix-sum / scatter passes.  Returns RBSP length in bytes (header bits + slice data + trailing
// bits, byte aligned).
long hh_cavlc_slice(const void *mbi, const uint8_t *nnz, const int16_t *coef, int mbw, int mbh, int srows, int slice,
                    int frame_i, uint64_t hdr_bits, int hdr_nbits, uint8_t *out, long cap)
{
    FrameSyntax fs{(const MbInfo *)mbi, nnz, coef, mbw, mbh, srows};
    const int per = slice_items_per(fs), nitems = mbw * mbh + (mbh + srows - 1) / srows;
    const int j0 = slice * per, j1 = per * (slice + 1) < nitems ? per * (slice + 1) : nitems;
    std::vector<unsigned long long> off(j1 - j0 + 1);
    unsigned long long pos = (unsigned long long)hdr_nbits;
    for (int j = j0; j < j1; j++) {
        SliceItem it = slice_item(fs, j);
        if (it.slice != slice || it.is_first != (j == j0) || it.is_end != (j == j1 - 1))
            return -3;
        BitCount c;
        int run = frame_i ? 0 : skip_run_before(fs, it.mb, it.first_mb);
        if (it.is_end)
            cavlc_end(c, run);
        else if (fs.mbi[it.mb].type != MB_PSKIP)
            cavlc_mb(c, fs, it.mb, frame_i, run);
        off[j - j0] = pos;
        pos += c.n;
    }
    long bytes = (long)((pos + 7) >> 3);
    if (bytes > cap)
        return -1;
    std::vector<uint32_t> buf((size_t)(bytes + 8) / 4 + 2, 0);
    {
        BitScatter s(buf.data(), 0);
        if (hdr_nbits > 32)
            s.put((uint32_t)(hdr_bits >> 32), hdr_nbits - 32);
        s.put((uint32_t)hdr_bits, hdr_nbits > 32 ? 32 : hdr_nbits);
        s.flush();
    }
    // scatter in reverse order to show the order does not matter
    for (int j = j1 - 1; j >= j0; j--) {
        SliceItem it = slice_item(fs, j);
        int run = frame_i ? 0 : skip_run_before(fs, it.mb, it.first_mb);
        if (!it.is_end && fs.mbi[it.mb].type == MB_PSKIP)
            continue;
        BitScatter s(buf.data(), off[j - j0]);
        if (it.is_end)
            cavlc_end(s, run);
        else
            cavlc_mb(s, fs, it.mb, frame_i, run);
        s.flush();
    }
    memcpy(out, buf.data(), (size_t)bytes);
    return bytes;
}

long hh_cabac_slice(const void *mbi, const uint8_t *nnz, const int16_t *coef, int mbw, int mbh, int srows, int slice,
                    int frame_i, int qp, uint64_t hdr_bits, int hdr_nbits, uint8_t *out, long cap)
{
    FrameSyntax fs{(const MbInfo *)mbi, nnz, coef, mbw, mbh, srows};
    const int row0 = slice * srows, row1 = row0 + srows < mbh ? row0 + srows : mbh;
    const int mb0 = row0 * mbw, nmb = (row1 - row0) * mbw;
    std::vector<size_t> off(nmb + 1);
    size_t total = 0;
    for (int i = 0; i < nmb; i++) {
        BinCount c;
        cabac_mb(c, fs, mb0 + i, frame_i);
        off[i] = total;
        total += c.n;
    }
    std::vector<uint16_t> bins(total + 1);
    for (int i = nmb - 1; i >= 0; i--) {
        BinWrite w(bins.data() + off[i]);
        cabac_mb(w, fs, mb0 + i, frame_i);
    }
    // header bits followed by cabac_alignment_one_bit up to the byte boundary
    int hb = (hdr_nbits + 7) >> 3;
    unsigned long long h = (hdr_bits << (hb * 8 - hdr_nbits)) | ((1ull << (hb * 8 - hdr_nbits)) - 1);
    for (int i = 0; i < hb; i++)
        out[i] = (uint8_t)(h >> (8 * (hb - 1 - i)));
    // pre-state resolution (sequential here; per-context parallel in cabac_kernel), then the coder
    CabacTables tab;
    tab.build(0, 1);
    uint32_t state[460];
    for (int i = 0; i < 460; i++)
        state[i] = cabac_init_state(i, frame_i, qp);
    std::vector<uint8_t> pre(total + 1, 0);
    for (size_t i = 0; i < total; i++) {
        uint16_t b = bins[i];
        if (!(b & (BIN_BYPASS | BIN_TERM))) {
            int ctx = b & 0x3ff;
            pre[i] = (uint8_t)state[ctx];
            state[ctx] = cabac_next_state(tab, state[ctx], (b >> 15) & 1);
        }
    }
    // stage 2 (range recurrence -> interval steps) and stage 3 (low recurrence -> bytes)
    CabacRange rc;
    std::vector<uint32_t> steps(total + 1);
    uint32_t flat_range = 510;
    for (size_t i = 0; i < total; i++) {
        uint32_t lps4 = tab.lpsw[(pre[i] >> 1) & 63], meta = cabac_stage_meta(bins[i], pre[i], tab);
        steps[i] = rc.step(lps4, meta);
        if (cabac_range_step_flat(flat_range, lps4, meta) != steps[i] || (flat_range != rc.range && i + 1 < total))
            return -2; // the branch-free variant the GPU uses must agree step by step
    }
    CabacBytes c;
    c.out = out + hb;
    for (size_t i = 0; i < total; i++) {
        if ((long)(hb + c.pos + 8) > cap)
            return -1;
        if (i + 1 < total)
            c.step_fast(steps[i]);
        else
            c.step(steps[i]);
    }
    return hb + (long)c.pos;
}

// Emulation prevention with the order-independent rule used by the GPU kernel.
long hh_epb(const uint8_t *in, long n, uint8_t *out, long cap)
{
    long o = 0;
    for (long i = 0; i < n; i++) {
        unsigned run = 0;
        for (long j = i - 1; j >= 0 && in[j] == 0; j--)
            run++;
        if (epb_needed(in[i], run)) {
            if (o >= cap)
                return -1;
            out[o++] = 3;
        }
        if (o >= cap)
            return -1;
        out[o++] = in[i];
    }
    return o;
}

int hh_sizeof_mbinfo() { return (int)sizeof(MbInfo); }

// Property the range stage relies on: on the MPS path the interval never needs more than a one-bit
// renormalisation (range - rangeTabLPS >= 128 for every state and every range of the quantiser cell).
int hh_mps_renorm_at_most_one()
{
    for (int s = 0; s < 64; s++)
        for (int range = 256; range <= 510; range++)
            if (range - (int)h264_range_lps[s][(range >> 6) & 3] < 128)
                return 0;
    return 1;
}
}
