/* Stand-in for <linux/clk.h>: see ../kstub.h (test infrastructure; not kernel code). */
#include "../kstub.h"
