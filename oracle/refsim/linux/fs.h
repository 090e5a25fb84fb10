/* Stand-in for <linux/fs.h>: see ../kstub.h (test infrastructure; not kernel code). */
#include "../kstub.h"
