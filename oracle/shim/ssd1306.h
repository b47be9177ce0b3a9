/* Stand-in for the lexus2k/ssd1306 header the reference includes at pi_sph_fluid.c:8.
 * Test infrastructure only: lets the reference translation unit compile headless into
 * oracle/_ref/ without the OLED driver tree.  Only the two symbols the reference calls
 * (pi_sph_fluid.c:468-469) are declared; both are no-ops defined in ssd1306_stub.c. */
#ifndef ORACLE_SHIM_SSD1306_H
#define ORACLE_SHIM_SSD1306_H
#include <stdint.h>
#include <unistd.h>
void ssd1306_128x64_i2c_init(void);
void ssd1306_drawBufferFast(int x, int y, int w, int h, const unsigned char *buf);
#endif
