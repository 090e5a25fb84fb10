/* Stand-in for <linux/of_device.h>: see ../kstub.h (test infrastructure; not kernel code). */
#include "../kstub.h"
