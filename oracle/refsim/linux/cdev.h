/* Stand-in for <linux/cdev.h>: see ../kstub.h (test infrastructure; not kernel code). */
#include "../kstub.h"
