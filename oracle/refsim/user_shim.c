/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  The libc side of user_shim.h: system-call semantics (-1 + errno) on top
 * of refsim.c's negative-errno entry points.
 */
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/ioctl.h>
#include <sys/mman.h>
#include <unistd.h>

int refsim_open(void);
long refsim_ioctl(unsigned int cmd, void *arg);
void *refsim_mmap(size_t size, uint32_t offset);
int refsim_release(void);

static int g_fd = -1;

static void release_at_exit(void) /* the kernel releases the device when the process exits (h264enc never closes it) */
{
    if (g_fd >= 0)
        refsim_release();
}

int refsim_u_open(const char *path, int flags, ...)
{
    va_list ap;
    va_start(ap, flags);
    int mode = (flags & O_CREAT) ? va_arg(ap, int) : 0;
    va_end(ap);
    if (strcmp(path, "/dev/cedar_dev"))
        return open(path, flags, mode);
    int r = refsim_open();
    if (r) {
        errno = -r;
        return -1;
    }
    g_fd = open("/dev/null", O_RDWR); /* reserves a descriptor number */
    atexit(release_at_exit);
    return g_fd;
}

int refsim_u_ioctl(int fd, unsigned long request, ...)
{
    va_list ap;
    va_start(ap, request);
    void *arg = va_arg(ap, void *);
    va_end(ap);
    if (fd != g_fd || g_fd < 0)
        return ioctl(fd, request, arg);
    long r = refsim_ioctl((unsigned int)request, arg);
    if (r < 0) {
        errno = (int)-r;
        return -1;
    }
    return (int)r;
}

void *refsim_u_mmap(void *addr, size_t length, int prot, int flags, int fd, off_t offset)
{
    if (fd != g_fd || g_fd < 0)
        return mmap(addr, length, prot, flags, fd, offset);
    void *p = refsim_mmap(length, (uint32_t)offset);
    if (!p) {
        errno = EINVAL;
        return MAP_FAILED;
    }
    return p;
}
