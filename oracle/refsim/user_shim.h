/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * user_shim.h is force-included (-include) when oracle/Makefile compiles the reference's own
 * userspace/h264enc.c, unmodified and from where it lies, into oracle/_ref/h264enc_sim: the program's
 * open / ioctl / mmap calls on /dev/cedar_dev are routed to the reference's kernel driver running in
 * the same process (refsim.c); every other file descriptor goes to libc as usual.
 */
#ifndef REFSIM_USER_SHIM_H
#define REFSIM_USER_SHIM_H
#include <fcntl.h>
#include <sys/ioctl.h>
#include <sys/mman.h>
#include <sys/types.h>
#include <unistd.h>

int refsim_u_open(const char *path, int flags, ...);
int refsim_u_ioctl(int fd, unsigned long request, ...);
void *refsim_u_mmap(void *addr, size_t length, int prot, int flags, int fd, off_t offset);

#define open refsim_u_open
#define ioctl refsim_u_ioctl
#define mmap refsim_u_mmap
#endif
