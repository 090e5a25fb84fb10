/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * kstub.h: a user-space stand-in for the handful of Linux kernel interfaces that
 * /root/reference/kernel/cedar.c uses, so that the reference's driver -- UNMODIFIED, compiled from the
 * file where it lies (oracle/Makefile, target refsim) -- runs here as ordinary user-space code against a
 * software model of the Cedar video engine's register file (ve_model.c).  Nothing in this file is kernel
 * code and nothing is copied from the reference; every definition is the minimum that makes the
 * reference's own source compile and behave as it would under the kernel:
 *
 *   readl / writel              -> ve_readl / ve_writel (the register model)
 *   dma_alloc_coherent          -> page-aligned host memory + a fake 32-bit bus address (kdma_*)
 *   copy_{from,to}_user         -> memcpy
 *   wait_event_..._timeout      -> the model raises the interrupt synchronously inside the trigger write
 *   clocks, resets, SRAM, cdev  -> succeed and do nothing
 *   dev_err / dev_info / pr_info-> a log that tests can read (stderr when REFSIM_VERBOSE is set)
 */
#ifndef REFSIM_KSTUB_H
#define REFSIM_KSTUB_H

#include <errno.h>
#include <stdarg.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h> /* dev_t, loff_t */
#include <time.h>

/* ---- annotations and small macros ---- */
#define __iomem
#define __user
#define __maybe_unused __attribute__((unused))
#define GFP_KERNEL 0
#define HZ 100
#define THIS_MODULE NULL
#define IORESOURCE_MEM 0x200
#define VM_IO 0x4000
#define IRQ_NONE 0
#define IRQ_HANDLED 1
#define ALIGN(x, a) (((x) + ((__typeof__(x))(a) - 1)) & ~((__typeof__(x))(a) - 1))
#define container_of(ptr, type, member) ((type *)((char *)(ptr) - offsetof(type, member)))
#define IS_ERR(p) ((uintptr_t)(p) >= (uintptr_t)-4095)
#define PTR_ERR(p) ((long)(intptr_t)(p))
#define of_match_ptr(p) (p)
#define MODULE_DEVICE_TABLE(a, b)
#define MODULE_DESCRIPTION(s)
#define MODULE_AUTHOR(s)
#define MODULE_LICENSE(s)
#define SET_RUNTIME_PM_OPS(suspend, resume, idle) .runtime_suspend = suspend, .runtime_resume = resume, .runtime_idle = idle
#define no_llseek NULL

/* ---- types ---- */
typedef uint32_t dma_addr_t; /* the A20 is a 32-bit ARM: bus addresses are 32 bits wide (the driver prints them with %08X) */
typedef uint32_t phys_addr_t;
typedef int irqreturn_t;
typedef unsigned long pgprot_t;
typedef irqreturn_t (*irq_handler_t)(int, void *);
typedef struct { int unused; } wait_queue_head_t;

struct device { void *drvdata; const char *name; };
struct module;
struct class { int unused; };
struct clk { int enabled; unsigned long rate; };
struct reset_control { int resets; };
struct resource { unsigned long start, end; };
struct file_operations;
struct cdev { struct module *owner; const struct file_operations *ops; };
struct inode { struct cdev *i_cdev; };
struct file { void *private_data; };
struct vm_area_struct {
    unsigned long vm_start, vm_end, vm_pgoff, vm_flags;
    pgprot_t vm_page_prot;
    /* filled by vm_iomap_memory(): what the mapping resolves to */
    phys_addr_t mapped_addr;
    size_t mapped_size;
};
struct file_operations {
    struct module *owner;
    int (*open)(struct inode *, struct file *);
    int (*release)(struct inode *, struct file *);
    loff_t (*llseek)(struct file *, loff_t, int);
    long (*unlocked_ioctl)(struct file *, unsigned int, unsigned long);
    int (*mmap)(struct file *, struct vm_area_struct *);
};
struct dev_pm_ops {
    int (*runtime_suspend)(struct device *);
    int (*runtime_resume)(struct device *);
    int (*runtime_idle)(struct device *);
};
struct of_device_id { const char *compatible; };
struct platform_device { struct device dev; };
struct device_driver {
    const char *name;
    const struct of_device_id *of_match_table;
    const struct dev_pm_ops *pm;
};
struct platform_driver {
    int (*probe)(struct platform_device *);
    int (*remove)(struct platform_device *);
    struct device_driver driver;
};
/* the reference ends with module_platform_driver(x): expose x to the harness */
#define module_platform_driver(drv) struct platform_driver *refsim_platform_driver = &(drv);

/* ---- the register model (ve_model.c) ---- */
uint32_t ve_readl(const volatile void *addr);
void ve_writel(uint32_t value, volatile void *addr);
void *ve_mmio_base(void);
void ve_attach_irq(irq_handler_t handler, void *dev_id);
#define readl(a) ve_readl(a)
#define writel(v, a) ve_writel((v), (a))

/* ---- logging ---- */
void klog(const char *level, const char *fmt, ...);
#define dev_err(dev, ...) klog("err", __VA_ARGS__)
#define dev_info(dev, ...) klog("info", __VA_ARGS__)
#define pr_info(...) klog("info", __VA_ARGS__)

/* ---- DMA memory: host pages with fake bus addresses ---- */
void *kdma_alloc(size_t size, dma_addr_t *handle);
void kdma_free(void *virt, dma_addr_t handle);
void *kdma_lookup(dma_addr_t addr, size_t *size_left); /* bus address -> host pointer (may point inside an allocation) */
static inline void *dma_alloc_coherent(struct device *dev, size_t size, dma_addr_t *handle, int gfp)
{
    (void)dev, (void)gfp;
    return kdma_alloc(size, handle);
}
static inline void dma_free_coherent(struct device *dev, size_t size, void *virt, dma_addr_t handle)
{
    (void)dev, (void)size;
    kdma_free(virt, handle);
}

/* ---- user access ---- */
static inline unsigned long copy_from_user(void *to, const void *from, unsigned long n) { memcpy(to, from, n); return 0; }
static inline unsigned long copy_to_user(void *to, const void *from, unsigned long n) { memcpy(to, from, n); return 0; }

/* ---- time, wait queues, interrupts ---- */
static inline uint64_t ktime_get_raw_fast_ns(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC_RAW, &ts);
    return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}
static inline void init_waitqueue_head(wait_queue_head_t *q) { (void)q; }
static inline void wake_up_interruptible(wait_queue_head_t *q) { (void)q; }
/* The model calls the interrupt handler from inside the trigger write, so the condition already holds (or the
 * encode "timed out": the model declined to run). */
#define wait_event_interruptible_timeout(q, cond, timeout) ((cond) ? 1 : 0)
static inline int devm_request_irq(struct device *dev, int irq, irq_handler_t handler, unsigned long flags, const char *name,
                                   void *dev_id)
{
    (void)dev, (void)irq, (void)flags, (void)name;
    ve_attach_irq(handler, dev_id);
    return 0;
}

/* ---- platform glue: everything succeeds ---- */
static inline int of_reserved_mem_device_init(struct device *dev) { (void)dev; return 0; }
static inline void of_reserved_mem_device_release(struct device *dev) { (void)dev; }
static inline int sunxi_sram_claim(struct device *dev) { (void)dev; return 0; }
static inline int sunxi_sram_release(struct device *dev) { (void)dev; return 0; }
static inline struct clk *devm_clk_get(struct device *dev, const char *id)
{
    (void)dev, (void)id;
    return (struct clk *)calloc(1, sizeof(struct clk));
}
static inline int clk_set_rate(struct clk *c, unsigned long rate) { c->rate = rate; return 0; }
static inline int clk_prepare_enable(struct clk *c) { c->enabled++; return 0; }
static inline void clk_disable_unprepare(struct clk *c) { c->enabled--; }
static inline struct reset_control *devm_reset_control_get(struct device *dev, const char *id)
{
    (void)dev, (void)id;
    return (struct reset_control *)calloc(1, sizeof(struct reset_control));
}
static inline int reset_control_reset(struct reset_control *r) { r->resets++; return 0; }
static inline int reset_control_assert(struct reset_control *r) { (void)r; return 0; }
static inline struct resource *platform_get_resource(struct platform_device *pdev, unsigned type, unsigned num)
{
    static struct resource res = {0x01c0e000, 0x01c0efff}; /* the A20's VE register window */
    (void)pdev, (void)type, (void)num;
    return &res;
}
static inline void *devm_ioremap_resource(struct device *dev, struct resource *res)
{
    (void)dev, (void)res;
    return ve_mmio_base();
}
static inline int platform_get_irq(struct platform_device *pdev, unsigned num) { (void)pdev, (void)num; return 85; }
static inline void *dev_get_drvdata(const struct device *dev) { return dev->drvdata; }
static inline void platform_set_drvdata(struct platform_device *pdev, void *data) { pdev->dev.drvdata = data; }
static inline void *platform_get_drvdata(const struct platform_device *pdev) { return pdev->dev.drvdata; }
static inline void *devm_kzalloc(struct device *dev, size_t size, int gfp) { (void)dev, (void)gfp; return calloc(1, size); }

/* ---- character device ---- */
static inline int alloc_chrdev_region(dev_t *dev, unsigned first, unsigned count, const char *name)
{
    (void)first, (void)count, (void)name;
    *dev = (240u << 20) | 0;
    return 0;
}
static inline void unregister_chrdev_region(dev_t dev, unsigned count) { (void)dev, (void)count; }
static inline void cdev_init(struct cdev *c, const struct file_operations *fops) { c->ops = fops; }
static inline int cdev_add(struct cdev *c, dev_t dev, unsigned count) { (void)c, (void)dev, (void)count; return 0; }
static inline void cdev_del(struct cdev *c) { (void)c; }
static inline struct class *class_create(struct module *owner, const char *name)
{
    (void)owner, (void)name;
    return (struct class *)calloc(1, sizeof(struct class));
}
static inline void class_destroy(struct class *c) { free(c); }
static inline struct device *device_create(struct class *c, struct device *parent, dev_t devt, void *drvdata, const char *fmt, ...)
{
    (void)c, (void)parent, (void)devt, (void)drvdata, (void)fmt;
    return (struct device *)calloc(1, sizeof(struct device));
}
static inline void device_destroy(struct class *c, dev_t devt) { (void)c, (void)devt; }

/* ---- mmap ---- */
static inline pgprot_t pgprot_noncached(pgprot_t p) { return p | 1; }
static inline int vm_iomap_memory(struct vm_area_struct *vma, phys_addr_t start, unsigned long len)
{
    vma->mapped_addr = start;
    vma->mapped_size = len;
    return 0;
}

#endif /* REFSIM_KSTUB_H */
