/* Stand-in for <linux/reset.h>: see ../kstub.h (test infrastructure; not kernel code). */
#include "../kstub.h"
