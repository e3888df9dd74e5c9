// Thread-local last-error slot shared by the translation units of libunetb200.so
// (read back through unetb200_last_error()); defined in unet_b200.cu.
#pragma once
__attribute__((visibility("hidden"))) int ub_fail(int code, const char* msg);
