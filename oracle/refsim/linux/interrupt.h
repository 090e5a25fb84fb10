/* Stand-in for <linux/interrupt.h>: see ../kstub.h (test infrastructure; not kernel code). */
#include "../kstub.h"
