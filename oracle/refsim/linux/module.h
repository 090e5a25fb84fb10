/* Stand-in for <linux/module.h>: see ../kstub.h (test infrastructure; not kernel code). */
#include "../kstub.h"
