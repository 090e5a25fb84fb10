/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * refsim.c: runs the reference's kernel driver in user space.  The driver source is NOT copied: this translation
 * unit #includes /root/reference/kernel/cedar.c from where it lies (oracle/Makefile passes -I$(REF)), after kstub.h
 * has stood in for the kernel headers it asks for (oracle/refsim/linux/...).  What is exported is a file-descriptor-free
 * version of the four system calls userspace/h264enc.c makes on /dev/cedar_dev:
 *
 *   refsim_open()            open("/dev/cedar_dev")        -> cedar_probe() once, then cedar_slashdev_open()
 *   refsim_ioctl(cmd, arg)   ioctl(fd, CEDAR_IOCTL_*, arg) -> cedar_slashdev_ioctl()  (negative errno, as the syscall layer sees it)
 *   refsim_mmap(size, off)   mmap(.., fd, off)             -> cedar_slashdev_mmap(), resolved to the host pages
 *   refsim_release()         close(fd)                     -> cedar_slashdev_release()
 *
 * plus read-only views of the driver's private state for the parity tests.
 */
#include "kstub.h"

#include "kernel/cedar.c" /* the reference, unmodified, from -I$(REF) */

void ve_sideband(int src_w, int src_h, int dst_w, int dst_h, int me_range);
void ve_reset_encoder(void);

static struct platform_device g_pdev;
static struct sunxi_cedar *g_cedar;
static struct inode g_inode;
static struct file g_file;
static int g_open;
static int g_me_range = 16;

void refsim_set_me_range(int r) { g_me_range = r > 0 ? r : 16; }

int refsim_open(void)
{
    if (g_open)
        return -EBUSY; /* one opener, like the chardev's single global context */
    if (!g_cedar) {
        g_pdev.dev.name = "cedar";
        int r = refsim_platform_driver->probe(&g_pdev);
        if (r)
            return r;
        g_cedar = (struct sunxi_cedar *)platform_get_drvdata(&g_pdev);
        g_inode.i_cdev = &g_cedar->cdev;
    }
    int r = g_cedar->cdev.ops->open(&g_inode, &g_file);
    if (r)
        return r;
    ve_reset_encoder();
    g_open = 1;
    return 0;
}

long refsim_ioctl(unsigned int cmd, void *arg)
{
    if (!g_open)
        return -EBADF;
    long r = g_cedar->cdev.ops->unlocked_ioctl(&g_file, cmd, (unsigned long)arg);
    if (cmd == CEDAR_IOCTL_CONFIG && r == 0) /* what the registers never carry (see ve_model.c) */
        ve_sideband(g_cedar->src_width, g_cedar->src_height, g_cedar->dst_width, g_cedar->dst_height, g_me_range);
    return r;
}

void *refsim_mmap(size_t size, uint32_t offset)
{
    if (!g_open)
        return NULL;
    struct vm_area_struct vma;
    memset(&vma, 0, sizeof vma);
    vma.vm_start = 0x10000000ul;
    vma.vm_end = vma.vm_start + size;
    vma.vm_pgoff = offset >> 12;
    if (g_cedar->cdev.ops->mmap(&g_file, &vma))
        return NULL;
    size_t left = 0;
    void *p = kdma_lookup(vma.mapped_addr, &left);
    return p && left >= vma.mapped_size ? p : NULL;
}

int refsim_release(void)
{
    if (!g_open)
        return -EBADF;
    g_open = 0;
    return g_cedar->cdev.ops->release(&g_inode, &g_file);
}

/* Driver-private state, by name (tests only). */
long refsim_state(const char *what)
{
    struct sunxi_cedar *c = g_cedar;
    if (!c)
        return -1;
#define F(name) if (!strcmp(what, #name)) return (long)c->name
    F(configured); F(src_width); F(src_height); F(src_format); F(src_width_mb); F(src_height_mb); F(src_stride_mb);
    F(dst_width); F(dst_height); F(dst_width_mb); F(dst_height_mb); F(dst_width_crop); F(dst_height_crop);
    F(profile); F(level); F(qp); F(keyframe_interval); F(frame_p_count); F(frame_count); F(entropy_coding_mode_cabac);
    F(input_luma_size); F(input_chroma_size); F(mb_info_size); F(mv_buffer_size); F(bytestream_size);
    F(thumbnail_enable); F(thumbnail_downscale);
#undef F
    if (!strcmp(what, "ref_luma_size")) return (long)c->reference_frame[0].luma_size;
    if (!strcmp(what, "ref_chroma_size")) return (long)c->reference_frame[0].chroma_size;
    if (!strcmp(what, "ref_subpic_size")) return (long)c->reference_frame[0].subpic_size;
    if (!strcmp(what, "reference_current")) return c->reference_current ? (long)(c->reference_current - c->reference_frame) : -1;
    return -2;
}
