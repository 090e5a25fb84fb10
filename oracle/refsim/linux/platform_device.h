/* Stand-in for <linux/platform_device.h>: see ../kstub.h (test infrastructure; not kernel code). */
#include "../kstub.h"
