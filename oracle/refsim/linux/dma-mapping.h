/* Stand-in for <linux/dma-mapping.h>: see ../kstub.h (test infrastructure; not kernel code). */
#include "../kstub.h"
