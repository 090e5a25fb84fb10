/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * ve_model.c: a software model of the part of the Allwinner Cedar video engine that the reference driver
 * (/root/reference/kernel/cedar.c, compiled unmodified through kstub.h) talks to.  Register offsets come from the
 * reference's own kernel/cedar_regs.h (included from where it lies, -I$(REF)/kernel); the behaviour is what the
 * driver's accesses imply:
 *
 *   PUTBITSDATA + STARTTRIG[3:0]=1, size in STARTTRIG[12:8]   append `size` bits MSB first   (cedar.c:187-207)
 *   INT_STATUS bit 9                                           put-bits engine idle            (cedar.c:195-199)
 *   PARA0 bit 31                                               emulation prevention OFF        (cedar.c:872-880)
 *   STMLEN                                                     bits in the stream so far       (cedar.c:886, 1208)
 *   STMOST / STMSTARTADDR / STMENDADDR / STMVSIZE              where the stream goes           (cedar.c:1052-1056)
 *   STARTTRIG = 0x08                                           encode one picture              (cedar.c:1176)
 *   INT_ENABLE / INT_STATUS bits 0..3, write-one-to-clear      completion interrupt            (cedar.c:225-249)
 *
 * The macroblock encoder behind the 0x08 trigger is silicon with no source anywhere; here it is the golden model
 * (oracle/h264_golden.c): the model reads the picture from the ISP input addresses, encodes it as the slice type PARA0
 * asks for at the QP PARA1 asks for, and appends the SLICE DATA (everything after the slice header the driver wrote
 * through the put-bits port) to the stream, emulation prevention applied by the stream unit as on the chip.
 * So: every header bit, every counter, every buffer size and every register value in a refsim run is the reference's
 * own code executing; only the slice data comes from the golden model.
 *
 * What the registers do NOT carry, and the model therefore receives out of band from refsim.c (ve_sideband): the exact
 * picture size in pixels (the driver programs macroblock counts only, cedar.c:1068-1074, and a stride of
 * 16 * src_stride_mb; the model reads rows packed at src_width as userspace/h264enc.c:178-187 writes them -- the two
 * agree whenever the width is a multiple of 16, which is what the README demands), the coded size (PICINFO is never
 * written) and the search range (MEPARA's fields are undocumented).  src_format never reaches a register at all
 * (cedar.c:793 stores it and nothing reads it): the model always reads NV12.
 */
#include "kstub.h"
#include "cedar_regs.h" /* the reference's register map, from $(REF)/kernel */
#include "../h264_golden.h"

#define VE_REGS_BYTES 0x1000
static uint32_t g_regs[VE_REGS_BYTES / 4];
static irq_handler_t g_irq;
static void *g_irq_dev;

/* ---- log ---- */
static char g_log[1 << 16];
static size_t g_log_len;
void klog(const char *level, const char *fmt, ...)
{
    char line[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(line, sizeof line, fmt, ap);
    va_end(ap);
    if (getenv("REFSIM_VERBOSE"))
        fprintf(stderr, "[cedar %s] %s", level, line);
    size_t n = strlen(line) + strlen(level) + 2;
    if (g_log_len + n + 1 >= sizeof g_log)
        g_log_len = 0; /* wrap: tests read the log right after the call they are interested in */
    g_log_len += (size_t)snprintf(g_log + g_log_len, sizeof g_log - g_log_len, "%s: %s", level, line);
}
const char *refsim_log(void) { return g_log; }
void refsim_log_clear(void) { g_log_len = 0, g_log[0] = 0; }

/* ---- DMA memory ---- */
#define KDMA_MAX 64
static struct { dma_addr_t addr; size_t size; void *virt; } g_dma[KDMA_MAX];
static dma_addr_t g_dma_next = 0x4A000000u;
void *kdma_alloc(size_t size, dma_addr_t *handle)
{
    size_t pages = (size + 4095) & ~(size_t)4095;
    for (int i = 0; i < KDMA_MAX; i++)
        if (!g_dma[i].virt) {
            void *p = NULL;
            if (posix_memalign(&p, 4096, pages ? pages : 4096))
                return NULL;
            memset(p, 0, pages);
            g_dma[i].virt = p, g_dma[i].size = pages, g_dma[i].addr = g_dma_next;
            g_dma_next += (dma_addr_t)pages + 4096; /* a guard page between allocations */
            *handle = g_dma[i].addr;
            return p;
        }
    return NULL;
}
void kdma_free(void *virt, dma_addr_t handle)
{
    for (int i = 0; i < KDMA_MAX; i++)
        if (g_dma[i].virt == virt && g_dma[i].addr == handle) {
            free(virt);
            g_dma[i].virt = NULL;
            return;
        }
    klog("err", "kdma_free(): unknown buffer 0x%08X\n", handle);
}
void *kdma_lookup(dma_addr_t addr, size_t *size_left)
{
    for (int i = 0; i < KDMA_MAX; i++)
        if (g_dma[i].virt && addr >= g_dma[i].addr && addr - g_dma[i].addr < g_dma[i].size) {
            if (size_left)
                *size_left = g_dma[i].size - (addr - g_dma[i].addr);
            return (uint8_t *)g_dma[i].virt + (addr - g_dma[i].addr);
        }
    return NULL;
}
int refsim_dma_live(void)
{
    int n = 0;
    for (int i = 0; i < KDMA_MAX; i++)
        n += g_dma[i].virt != NULL;
    return n;
}

/* ---- stream unit ---- */
static struct {
    uint64_t acc; /* pending bits, right aligned */
    int nacc;
    size_t bytes; /* bytes written to the stream buffer */
    int zeros;    /* run of 0x00 bytes, for emulation prevention */
    int overflow;
} g_stm;

static void stm_byte(uint8_t b)
{
    size_t left = 0;
    uint8_t *base = (uint8_t *)kdma_lookup(g_regs[(CEDAR_H264ENC_BASE + CEDAR_H264ENC_STMSTARTADDR) / 4], &left);
    size_t cap = (size_t)g_regs[(CEDAR_H264ENC_BASE + CEDAR_H264ENC_STMENDADDR) / 4] -
                 g_regs[(CEDAR_H264ENC_BASE + CEDAR_H264ENC_STMSTARTADDR) / 4] + 1;
    if (!base || g_stm.bytes >= cap || g_stm.bytes >= left) {
        g_stm.overflow = 1;
        return;
    }
    base[g_stm.bytes++] = b;
}

static void stm_put(uint32_t data, int size)
{
    const int epb_off = (g_regs[(CEDAR_H264ENC_BASE + CEDAR_H264ENC_PARA0) / 4] >> 31) & 1;
    if (size <= 0)
        return;
    g_stm.acc = (g_stm.acc << size) | (data & (size >= 32 ? 0xFFFFFFFFu : ((1u << size) - 1)));
    g_stm.nacc += size;
    while (g_stm.nacc >= 8) {
        uint8_t b = (uint8_t)(g_stm.acc >> (g_stm.nacc - 8));
        g_stm.nacc -= 8;
        if (epb_off) {
            stm_byte(b);
            g_stm.zeros = 0; /* start code + NAL header: the unit's payload starts after them */
        } else {
            if (g_stm.zeros >= 2 && b <= 3) {
                stm_byte(0x03);
                g_stm.zeros = 0;
            }
            stm_byte(b);
            g_stm.zeros = b == 0 ? g_stm.zeros + 1 : 0;
        }
    }
}

/* ---- the "silicon" ---- */
static struct {
    int src_w, src_h, dst_w, dst_h, me_range;
} g_side;
static gm_encoder *g_gm;
static gm_config g_gm_cfg;
static uint32_t g_prev_rec_y;
static int g_pictures;
static char g_fault[256];

void ve_sideband(int src_w, int src_h, int dst_w, int dst_h, int me_range)
{
    g_side.src_w = src_w, g_side.src_h = src_h, g_side.dst_w = dst_w, g_side.dst_h = dst_h, g_side.me_range = me_range;
}
const char *refsim_ve_fault(void) { return g_fault; }
int refsim_ve_pictures(void) { return g_pictures; }

void ve_reset_encoder(void)
{
    if (g_gm)
        gm_close(g_gm);
    g_gm = NULL;
    g_prev_rec_y = 0;
    g_fault[0] = 0;
}

#define ENC(r) g_regs[(CEDAR_H264ENC_BASE + (r)) / 4]
#define ISP(r) g_regs[(CEDAR_H264ISP_BASE + (r)) / 4]
#define FAULT(...) (snprintf(g_fault, sizeof g_fault, __VA_ARGS__), klog("err", "VE model: %s\n", g_fault), -1)

static int ve_encode_picture(void)
{
    const uint32_t para0 = ENC(CEDAR_H264ENC_PARA0), para1 = ENC(CEDAR_H264ENC_PARA1);
    const int slice_type = (para0 >> 4) & 7, cabac = (para0 >> 8) & 1, qp = para1 & 0xFF;
    const int frame_i = slice_type == 0;
    if (slice_type > 1)
        return FAULT("PARA0 slice type %d: B pictures are not modelled", slice_type);
    if (((para1 >> 8) & 0xFF) != (uint32_t)qp)
        return FAULT("PARA1: fixed QP %d != fixed intra QP %d", qp, (para1 >> 8) & 0xFF);
    if (!g_side.src_w)
        return FAULT("no picture geometry");
    /* macroblock counts the driver programmed must describe the picture it was configured with */
    if (ISP(CEDAR_H264ISP_INPUT_SIZE) != ((uint32_t)((g_side.src_w + 15) >> 4) << 16 | (uint32_t)((g_side.src_h + 15) >> 4)))
        return FAULT("ISP_INPUT_SIZE 0x%08X does not match %dx%d", ISP(CEDAR_H264ISP_INPUT_SIZE), g_side.src_w, g_side.src_h);
    if (!frame_i && ENC(CEDAR_H264ENC_REFADDRY) != g_prev_rec_y)
        return FAULT("P picture references 0x%08X, the previous picture was reconstructed at 0x%08X",
                     ENC(CEDAR_H264ENC_REFADDRY), g_prev_rec_y);
    if (ENC(CEDAR_H264ENC_RECADDRY) == g_prev_rec_y && g_prev_rec_y)
        return FAULT("reconstruction would overwrite its own reference");
    size_t luma_left = 0, chroma_left = 0;
    const uint8_t *luma = (const uint8_t *)kdma_lookup(ISP(CEDAR_H264ISP_INPUT_Y_ADDR), &luma_left);
    const uint8_t *chroma = (const uint8_t *)kdma_lookup(ISP(CEDAR_H264ISP_INPUT_C0_ADDR), &chroma_left);
    const size_t ysz = (size_t)g_side.src_w * g_side.src_h;
    if (!luma || !chroma || luma_left < ysz || chroma_left < ysz / 2)
        return FAULT("input buffers not mapped or too small");

    gm_config want;
    memset(&want, 0, sizeof want);
    want.src_width = g_side.src_w, want.src_height = g_side.src_h, want.src_format = GM_FORMAT_NV12;
    want.dst_width = g_side.dst_w, want.dst_height = g_side.dst_h;
    want.profile = 77, want.level = 41; /* only used by the parameter sets, which are not taken from the model */
    want.qp = qp, want.keyframe_interval = 1 << 30, want.relax_gop = 1; /* the picture type comes from PARA0 */
    want.entropy_coding_mode = cabac ? GM_ENTROPY_CABAC : GM_ENTROPY_CAVLC;
    want.me_range = g_side.me_range;
    if (!g_gm || memcmp(&want, &g_gm_cfg, sizeof want)) {
        if (g_gm)
            gm_close(g_gm);
        g_gm = NULL;
        if (gm_open(&want, &g_gm))
            return FAULT("golden model refused the configuration");
        g_gm_cfg = want;
        if (!frame_i)
            return FAULT("P picture without a reference");
    }
    gm_set_frame_p_count(g_gm, frame_i ? 0 : 1);
    const int cap = g_side.dst_w * g_side.dst_h * 3 + 65536;
    uint8_t *tmp = (uint8_t *)malloc((size_t)cap * 2);
    if (!tmp)
        return FAULT("out of memory");
    int n = gm_encode_frame(g_gm, luma, chroma, tmp, cap);
    if (n < 0) {
        free(tmp);
        return FAULT("golden model failed: %d", n);
    }
    /* find the slice NAL (the model may have put parameter sets in front of its first picture) */
    int pos = -1;
    for (int i = 0; i + 4 < n; i++)
        if (!tmp[i] && !tmp[i + 1] && !tmp[i + 2] && tmp[i + 3] == 1 && ((tmp[i + 4] & 0x1F) == 5 || (tmp[i + 4] & 0x1F) == 1)) {
            pos = i + 5;
            break;
        }
    if (pos < 0) {
        free(tmp);
        return FAULT("no slice NAL in the golden model's output");
    }
    /* undo emulation prevention -> RBSP */
    uint8_t *rbsp = tmp + cap;
    int m = 0, zeros = 0;
    for (int i = pos; i < n; i++) {
        if (zeros >= 2 && tmp[i] == 3) {
            zeros = 0;
            continue;
        }
        rbsp[m++] = tmp[i];
        zeros = tmp[i] == 0 ? zeros + 1 : 0;
    }
    /* skip the slice header (the driver wrote its own through the put-bits port) and append the rest bit-exactly */
    uint32_t hdr = 0;
    int hbits = 0;
    gm_slice_header_bits(frame_i, frame_i ? 0 : 1, cabac, &hdr, &hbits);
    for (long bit = hbits; bit < (long)m * 8;) {
        int take = 8 - (int)(bit & 7);
        uint32_t v = rbsp[bit >> 3] & ((1u << take) - 1);
        stm_put(v, take);
        bit += take;
    }
    free(tmp);
    if (g_stm.nacc)
        return FAULT("stream not byte aligned after the picture (%d bits pending): the driver's slice header has a "
                     "different length than the model's", g_stm.nacc);
    if (g_stm.overflow)
        return FAULT("bytestream buffer overflow");
    g_prev_rec_y = ENC(CEDAR_H264ENC_RECADDRY);
    g_pictures++;
    return 0;
}

/* ---- register file ---- */
void *ve_mmio_base(void) { return g_regs; }
void ve_attach_irq(irq_handler_t handler, void *dev_id) { g_irq = handler, g_irq_dev = dev_id; }
uint32_t refsim_reg(int offset) { return offset >= 0 && offset < VE_REGS_BYTES ? g_regs[offset / 4] : 0; }

uint32_t ve_readl(const volatile void *addr)
{
    const long off = (const uint8_t *)addr - (const uint8_t *)g_regs;
    if (off < 0 || off >= VE_REGS_BYTES) {
        klog("err", "VE model: read outside the register window (%ld)\n", off);
        return 0;
    }
    switch (off) {
    case CEDAR_VE_VERSION: return 0x1623u << 16; /* the A20's engine */
    case CEDAR_H264ENC_BASE + CEDAR_H264ENC_INT_STATUS: return g_regs[off / 4] | 0x200; /* put-bits engine idle */
    case CEDAR_H264ENC_BASE + CEDAR_H264ENC_STMLEN: return (uint32_t)(g_stm.bytes * 8 + (size_t)g_stm.nacc);
    default: return g_regs[off / 4];
    }
}

void ve_writel(uint32_t value, volatile void *addr)
{
    const long off = (const uint8_t *)addr - (const uint8_t *)g_regs;
    if (off < 0 || off >= VE_REGS_BYTES) {
        klog("err", "VE model: write outside the register window (%ld)\n", off);
        return;
    }
    switch (off) {
    case CEDAR_H264ENC_BASE + CEDAR_H264ENC_INT_STATUS: /* write one to clear */
        g_regs[off / 4] &= ~(value & 0xF);
        return;
    case CEDAR_H264ENC_BASE + CEDAR_H264ENC_STMOST: /* stream offset: the driver restarts the buffer with 0 per picture */
        g_regs[off / 4] = value;
        memset(&g_stm, 0, sizeof g_stm);
        g_stm.bytes = value / 8;
        return;
    case CEDAR_H264ENC_BASE + CEDAR_H264ENC_STARTTRIG:
        g_regs[off / 4] = value;
        if ((value & 0xF) == 0x1) {
            stm_put(ENC(CEDAR_H264ENC_PUTBITSDATA), (int)((value >> 8) & 0x1F));
        } else if ((value & 0xF) == 0x8) {
            int r = ve_encode_picture();
            ENC(CEDAR_H264ENC_INT_STATUS) |= r == 0 ? 0x1 : 0x2; /* done / error */
            if (g_irq && (ENC(CEDAR_H264ENC_INT_ENABLE) & 0x7))
                g_irq(85, g_irq_dev);
        }
        return;
    default:
        g_regs[off / 4] = value;
    }
}
