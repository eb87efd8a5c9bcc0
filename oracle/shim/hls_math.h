/* Empty stand-in, see ap_fixed.h in this directory. */
#pragma once
