/* Stand-in for <linux/mm.h>: see ../kstub.h (test infrastructure; not kernel code). */
#include "../kstub.h"
