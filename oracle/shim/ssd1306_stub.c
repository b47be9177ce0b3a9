/* No-op display functions (see shim/ssd1306.h). Test infrastructure only. */
#include "ssd1306.h"
void ssd1306_128x64_i2c_init(void) {}
void ssd1306_drawBufferFast(int x, int y, int w, int h, const unsigned char *buf) {
    (void)x; (void)y; (void)w; (void)h; (void)buf;
}
