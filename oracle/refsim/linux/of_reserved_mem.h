/* Stand-in for <linux/of_reserved_mem.h>: see ../kstub.h (test infrastructure; not kernel code). */
#include "../kstub.h"
