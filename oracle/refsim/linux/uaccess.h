/* Stand-in for <linux/uaccess.h>: see ../kstub.h (test infrastructure; not kernel code). */
#include "../kstub.h"
