/* Stand-in for <linux/soc/sunxi/sunxi_sram.h>: see kstub.h (test infrastructure). */
#include "../../../kstub.h"
