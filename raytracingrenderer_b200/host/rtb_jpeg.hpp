// rtb_jpeg.hpp — JPEG decoder for the host scene loader (textures of the bathroom scene).
//
// Baseline / extended-sequential and progressive Huffman JPEG, 8-bit, 1 / 3 / 4 components,
// restart intervals, JFIF / Adobe colour hints.  Entropy decoding follows ITU-T T.81.  The steps T.81
// leaves to the implementation are done with the arithmetic stb_image uses, because the reference
// decodes its textures with stb_image (RTBase/Imaging.h:32-71) and the flattened scene must match it
// bit for bit (tests/test_host_cpu.py):
//   * inverse DCT: the jidctint-derived integer transform with 12-bit constants, 2 extra bits kept
//     after the column pass, +128 level shift folded into the row pass rounding;
//   * chroma up-sampling: 3:1 triangle filters (h2, v2, h2v2 with /16 rounding), replication otherwise;
//   * YCbCr -> RGB: 20-bit fixed point with the constants rounded to 12 bits.
#pragma once

#include <cstdint>
#include <cstring>
#include <vector>

namespace rtb_img
{

class JpegDecoder
{
public:
	// -> interleaved 8-bit pixels, `channels` = 3 (colour) or 1 (grey).  false on any error.
	bool decode(const std::vector<unsigned char>& file, int& width, int& height, int& channels, std::vector<unsigned char>& out)
	{
		p = file.data();
		end = p + file.size();
		if (get8() != 0xFF || get8() != 0xD8) return false; // SOI
		// tables / misc until the frame header
		int m = nextMarker();
		while (!isSOF(m))
		{
			if (!tablesOrMisc(m)) return false;
			m = nextMarker();
			while (m == 0xFF)
			{
				if (p >= end) return false;
				m = nextMarker();
			}
		}
		progressive = (m == 0xC2);
		if (!frameHeader()) return false;
		// scans
		m = nextMarker();
		for (;;)
		{
			if (m == 0xDA)
			{
				if (!scanHeader()) return false;
				if (!entropyCodedData()) return false;
				if (pendingMarker == 0xFF)
				{
					// skip to the next marker
					while (p < end)
					{
						if (get8() == 0xFF)
						{
							pendingMarker = get8();
							break;
						}
					}
				}
				m = nextMarker();
				if (m >= 0xD0 && m <= 0xD7) m = nextMarker();
			}
			else if (m == 0xD9) break; // EOI
			else if (m == 0xDC)         // DNL
			{
				int len = get16();
				int lines = get16();
				if (len != 4 || lines != imgH) return false;
				m = nextMarker();
			}
			else
			{
				if (m == 0xFF && p >= end) break; // truncated file: keep what was decoded
				if (!tablesOrMisc(m)) return false;
				m = nextMarker();
			}
		}
		if (progressive) finishProgressive();
		return output(width, height, channels, out);
	}

private:
	struct Huffman
	{
		uint8_t values[256];
		int mincode[17], maxcode[18], valptr[17]; // per code length (T.81 F.2.2.3)
		bool present = false;
	};
	struct Component
	{
		int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0, dcPred = 0;
		int x = 0, y = 0, w2 = 0, h2 = 0, coeffW = 0, coeffH = 0;
		std::vector<uint8_t> data;
		std::vector<int16_t> coeff;
	};

	const unsigned char* p = nullptr;
	const unsigned char* end = nullptr;
	uint16_t dequant[4][64];
	Huffman dcTab[4], acTab[4];
	Component comp[4];
	int nComp = 0, imgW = 0, imgH = 0, hMax = 1, vMax = 1, mcuW = 0, mcuH = 0, mcusX = 0, mcusY = 0;
	bool progressive = false, jfif = false;
	int adobeTransform = -1, rgbIds = 0;
	int restartInterval = 0, todo = 0;
	int scanN = 0, order[4] = {0, 0, 0, 0};
	int specStart = 0, specEnd = 0, succHigh = 0, succLow = 0, eobRun = 0;
	// bit reader
	uint32_t bitBuf = 0;
	int bitCnt = 0;
	int pendingMarker = 0xFF; // 0xFF = none
	bool noMore = false;

	static const uint8_t* zigzag()
	{
		static const uint8_t z[64 + 15] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33,
		                                   40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36,
		                                   29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54,
		                                   47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};
		return z;
	}
	int get8() { return p < end ? *p++ : 0; }
	int get16()
	{
		int a = get8();
		return (a << 8) | get8();
	}
	static bool isSOF(int m) { return m == 0xC0 || m == 0xC1 || m == 0xC2; }
	int nextMarker()
	{
		if (pendingMarker != 0xFF)
		{
			int m = pendingMarker;
			pendingMarker = 0xFF;
			return m;
		}
		int x = get8();
		if (x != 0xFF) return 0xFF;
		while (x == 0xFF) x = get8(); // fill bytes
		return x;
	}

	bool tablesOrMisc(int m)
	{
		switch (m)
		{
		case 0xFF: return false; // expected a marker
		case 0xDD:               // DRI
			if (get16() != 4) return false;
			restartInterval = get16();
			return true;
		case 0xDB: // DQT
		{
			int len = get16() - 2;
			while (len > 0)
			{
				int q = get8();
				int prec = q >> 4, t = q & 15;
				if ((prec != 0 && prec != 1) || t > 3) return false;
				for (int i = 0; i < 64; i++) dequant[t][zigzag()[i]] = (uint16_t)(prec ? get16() : get8());
				len -= prec ? 129 : 65;
			}
			return len == 0;
		}
		case 0xC4: // DHT
		{
			int len = get16() - 2;
			while (len > 0)
			{
				int q = get8();
				int tc = q >> 4, th = q & 15;
				if (tc > 1 || th > 3) return false;
				int counts[17], total = 0;
				for (int i = 1; i <= 16; i++) total += (counts[i] = get8());
				if (total > 256) return false;
				Huffman& h = tc ? acTab[th] : dcTab[th];
				for (int i = 0; i < total; i++) h.values[i] = (uint8_t)get8();
				int code = 0, k = 0;
				for (int l = 1; l <= 16; l++)
				{
					h.valptr[l] = k;
					h.mincode[l] = code;
					code += counts[l];
					k += counts[l];
					h.maxcode[l] = counts[l] ? code - 1 : -1;
					code <<= 1;
				}
				h.maxcode[17] = 0x7FFFFFFF;
				h.present = true;
				len -= 17 + total;
			}
			return len == 0;
		}
		default: break;
		}
		if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE)
		{
			int len = get16();
			if (len < 2) return false;
			len -= 2;
			if (m == 0xE0 && len >= 5)
			{
				static const char tag[5] = {'J', 'F', 'I', 'F', 0};
				bool ok = true;
				for (int i = 0; i < 5; i++)
					if (get8() != (unsigned char)tag[i]) ok = false;
				len -= 5;
				if (ok) jfif = true;
			}
			else if (m == 0xEE && len >= 12)
			{
				static const char tag[6] = {'A', 'd', 'o', 'b', 'e', 0};
				bool ok = true;
				for (int i = 0; i < 6; i++)
					if (get8() != (unsigned char)tag[i]) ok = false;
				len -= 6;
				if (ok)
				{
					get8();
					get16();
					get16();
					adobeTransform = get8();
					len -= 6;
				}
			}
			if (p + len > end) return false;
			p += len;
			return true;
		}
		return false;
	}

	bool frameHeader()
	{
		int len = get16();
		if (get8() != 8) return false; // 8-bit samples only
		imgH = get16();
		imgW = get16();
		nComp = get8();
		if (imgH <= 0 || imgW <= 0 || (nComp != 1 && nComp != 3 && nComp != 4) || len != 8 + 3 * nComp) return false;
		static const char rgb[3] = {'R', 'G', 'B'};
		rgbIds = 0;
		hMax = vMax = 1;
		for (int i = 0; i < nComp; i++)
		{
			comp[i].id = get8();
			if (nComp == 3 && comp[i].id == rgb[i]) rgbIds++;
			int q = get8();
			comp[i].h = q >> 4, comp[i].v = q & 15;
			comp[i].tq = get8();
			if (comp[i].h < 1 || comp[i].h > 4 || comp[i].v < 1 || comp[i].v > 4 || comp[i].tq > 3) return false;
			if (comp[i].h > hMax) hMax = comp[i].h;
			if (comp[i].v > vMax) vMax = comp[i].v;
		}
		for (int i = 0; i < nComp; i++)
			if (hMax % comp[i].h || vMax % comp[i].v) return false;
		mcuW = hMax * 8, mcuH = vMax * 8;
		mcusX = (imgW + mcuW - 1) / mcuW, mcusY = (imgH + mcuH - 1) / mcuH;
		for (int i = 0; i < nComp; i++)
		{
			Component& c = comp[i];
			c.x = (imgW * c.h + hMax - 1) / hMax;
			c.y = (imgH * c.v + vMax - 1) / vMax;
			c.w2 = mcusX * c.h * 8;
			c.h2 = mcusY * c.v * 8;
			c.data.assign((size_t)c.w2 * c.h2, 0);
			if (progressive)
			{
				c.coeffW = c.w2 / 8, c.coeffH = c.h2 / 8;
				c.coeff.assign((size_t)c.w2 * c.h2, 0);
			}
		}
		return true;
	}

	bool scanHeader()
	{
		int len = get16();
		scanN = get8();
		if (scanN < 1 || scanN > 4 || scanN > nComp || len != 6 + 2 * scanN) return false;
		for (int i = 0; i < scanN; i++)
		{
			int id = get8(), q = get8(), which = -1;
			for (int k = 0; k < nComp; k++)
				if (comp[k].id == id) which = k;
			if (which < 0) return false;
			comp[which].hd = q >> 4, comp[which].ha = q & 15;
			if (comp[which].hd > 3 || comp[which].ha > 3) return false;
			order[i] = which;
		}
		specStart = get8();
		specEnd = get8();
		int a = get8();
		succHigh = a >> 4, succLow = a & 15;
		if (progressive)
		{
			if (specStart > 63 || specEnd > 63 || specStart > specEnd || succHigh > 13 || succLow > 13) return false;
		}
		else
		{
			if (specStart != 0 || succHigh != 0 || succLow != 0) return false;
			specEnd = 63;
		}
		return true;
	}

	// ---- bit reader: bytes are appended below the valid bits of a 32-bit window; 0xFF00 is a
	// stuffed 0xFF, any other 0xFFxx is a marker after which zeros are fed.
	void fill()
	{
		do
		{
			unsigned b = noMore ? 0 : (unsigned)get8();
			if (b == 0xFF)
			{
				int c = get8();
				while (c == 0xFF) c = get8();
				if (c != 0)
				{
					pendingMarker = c;
					noMore = true;
					return;
				}
			}
			bitBuf |= b << (24 - bitCnt);
			bitCnt += 8;
		} while (bitCnt <= 24);
	}
	int getBits(int n)
	{
		if (n == 0) return 0;
		if (bitCnt < n) fill();
		uint32_t v = bitBuf >> (32 - n);
		bitBuf <<= n;
		bitCnt -= n;
		if (bitCnt < 0) bitCnt = 0; // ran past a marker: zeros
		return (int)v;
	}
	int getBit() { return getBits(1); }
	int decodeSymbol(const Huffman& h)
	{
		if (bitCnt < 16) fill();
		int code = 0;
		for (int l = 1; l <= 16; l++)
		{
			code = (code << 1) | (int)(bitBuf >> 31);
			bitBuf <<= 1;
			bitCnt--;
			if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l])
			{
				if (bitCnt < 0) bitCnt = 0;
				return h.values[h.valptr[l] + code - h.mincode[l]];
			}
		}
		if (bitCnt < 0) bitCnt = 0;
		return -1;
	}
	// T.81 F.2.2.1 EXTEND of an n-bit magnitude
	int receiveExtend(int n)
	{
		if (n == 0) return 0;
		int v = getBits(n);
		return (v < (1 << (n - 1))) ? v - (1 << n) + 1 : v;
	}
	void resetDecoder()
	{
		bitBuf = 0, bitCnt = 0;
		noMore = false;
		pendingMarker = 0xFF;
		for (int i = 0; i < 4; i++) comp[i].dcPred = 0;
		todo = restartInterval ? restartInterval : 0x7FFFFFFF;
		eobRun = 0;
	}
	// after an MCU: restart-interval bookkeeping.  false = stop the scan (no restart marker where one is due)
	bool mcuDone()
	{
		if (--todo <= 0)
		{
			if (bitCnt < 24) fill();
			if (!(pendingMarker >= 0xD0 && pendingMarker <= 0xD7)) return false;
			resetDecoder();
		}
		return true;
	}

	bool blockBaseline(int16_t* data, Component& c)
	{
		const Huffman& hdc = dcTab[c.hd];
		const Huffman& hac = acTab[c.ha];
		const uint16_t* dq = dequant[c.tq];
		int t = decodeSymbol(hdc);
		if (t < 0 || t > 15) return false;
		memset(data, 0, 64 * sizeof(int16_t));
		int diff = receiveExtend(t);
		c.dcPred += diff;
		data[0] = (int16_t)(c.dcPred * dq[0]);
		int k = 1;
		do
		{
			int rs = decodeSymbol(hac);
			if (rs < 0) return false;
			int s = rs & 15, r = rs >> 4;
			if (s == 0)
			{
				if (rs != 0xF0) break; // end of block
				k += 16;
			}
			else
			{
				k += r;
				int z = zigzag()[k++];
				data[z] = (int16_t)(receiveExtend(s) * dq[z]);
			}
		} while (k < 64);
		return true;
	}

	bool blockProgressiveDC(int16_t* data, Component& c)
	{
		if (specEnd != 0) return false;
		if (succHigh == 0)
		{
			memset(data, 0, 64 * sizeof(int16_t));
			int t = decodeSymbol(dcTab[c.hd]);
			if (t < 0 || t > 15) return false;
			c.dcPred += receiveExtend(t);
			data[0] = (int16_t)(c.dcPred * (1 << succLow));
		}
		else if (getBit())
			data[0] += (int16_t)(1 << succLow);
		return true;
	}

	bool blockProgressiveAC(int16_t* data, Component& c)
	{
		if (specStart == 0) return false;
		const Huffman& hac = acTab[c.ha];
		if (succHigh == 0)
		{
			if (eobRun)
			{
				--eobRun;
				return true;
			}
			int k = specStart;
			do
			{
				int rs = decodeSymbol(hac);
				if (rs < 0) return false;
				int s = rs & 15, r = rs >> 4;
				if (s == 0)
				{
					if (r < 15)
					{
						eobRun = (1 << r);
						if (r) eobRun += getBits(r);
						--eobRun;
						break;
					}
					k += 16;
				}
				else
				{
					k += r;
					int z = zigzag()[k++];
					data[z] = (int16_t)(receiveExtend(s) * (1 << succLow));
				}
			} while (k <= specEnd);
			return true;
		}
		// refinement of already-coded coefficients
		int16_t bit = (int16_t)(1 << succLow);
		auto refine = [&](int16_t* q) {
			if (getBit())
				if ((*q & bit) == 0) *q += (*q > 0) ? bit : (int16_t)-bit;
		};
		if (eobRun)
		{
			--eobRun;
			for (int k = specStart; k <= specEnd; k++)
			{
				int16_t* q = &data[zigzag()[k]];
				if (*q != 0) refine(q);
			}
			return true;
		}
		int k = specStart;
		do
		{
			int rs = decodeSymbol(hac);
			if (rs < 0) return false;
			int s = rs & 15, r = rs >> 4;
			if (s == 0)
			{
				if (r < 15)
				{
					eobRun = (1 << r) - 1;
					if (r) eobRun += getBits(r);
					r = 64; // run to the end of the band
				}
			}
			else
			{
				if (s != 1) return false;
				s = getBit() ? bit : -bit;
			}
			while (k <= specEnd)
			{
				int16_t* q = &data[zigzag()[k++]];
				if (*q != 0) refine(q);
				else
				{
					if (r == 0)
					{
						*q = (int16_t)s;
						break;
					}
					--r;
				}
			}
		} while (k <= specEnd);
		return true;
	}

	bool entropyCodedData()
	{
		resetDecoder();
		int16_t block[64];
		if (scanN == 1)
		{
			// non-interleaved: every block of the component in raster order is an MCU
			Component& c = comp[order[0]];
			int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
			for (int j = 0; j < h; j++)
				for (int i = 0; i < w; i++)
				{
					if (!progressive)
					{
						if (!blockBaseline(block, c)) return false;
						idct(&c.data[(size_t)c.w2 * j * 8 + i * 8], c.w2, block);
					}
					else
					{
						int16_t* d = &c.coeff[64 * ((size_t)i + (size_t)j * c.coeffW)];
						if (specStart == 0 ? !blockProgressiveDC(d, c) : !blockProgressiveAC(d, c)) return false;
					}
					if (!mcuDone()) return true;
				}
			return true;
		}
		for (int j = 0; j < mcusY; j++)
			for (int i = 0; i < mcusX; i++)
			{
				for (int k = 0; k < scanN; k++)
				{
					Component& c = comp[order[k]];
					for (int y = 0; y < c.v; y++)
						for (int x = 0; x < c.h; x++)
						{
							int bx = i * c.h + x, by = j * c.v + y;
							if (!progressive)
							{
								if (!blockBaseline(block, c)) return false;
								idct(&c.data[(size_t)c.w2 * by * 8 + bx * 8], c.w2, block);
							}
							else
							{
								// interleaved progressive scans carry DC only
								if (!blockProgressiveDC(&c.coeff[64 * ((size_t)bx + (size_t)by * c.coeffW)], c)) return false;
							}
						}
				}
				if (!mcuDone()) return true;
			}
		return true;
	}

	void finishProgressive()
	{
		for (int n = 0; n < nComp; n++)
		{
			Component& c = comp[n];
			int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
			for (int j = 0; j < h; j++)
				for (int i = 0; i < w; i++)
				{
					int16_t* d = &c.coeff[64 * ((size_t)i + (size_t)j * c.coeffW)];
					for (int k = 0; k < 64; k++) d[k] = (int16_t)(d[k] * dequant[c.tq][k]);
					idct(&c.data[(size_t)c.w2 * j * 8 + i * 8], c.w2, d);
				}
		}
	}

	// ---- inverse DCT (integer, 12-bit constants)
	static int fix(double x) { return (int)(x * 4096 + 0.5); }
	struct Idct1D
	{
		int x0, x1, x2, x3, t0, t1, t2, t3;
	};
	static Idct1D idct1d(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7)
	{
		Idct1D r;
		int p1, p2, p3, p4, p5, t0, t1, t2, t3;
		p2 = s2, p3 = s6;
		p1 = (p2 + p3) * fix(0.5411961f);
		t2 = p1 + p3 * fix(-1.847759065f);
		t3 = p1 + p2 * fix(0.765366865f);
		p2 = s0, p3 = s4;
		t0 = (p2 + p3) * 4096;
		t1 = (p2 - p3) * 4096;
		r.x0 = t0 + t3, r.x3 = t0 - t3, r.x1 = t1 + t2, r.x2 = t1 - t2;
		t0 = s7, t1 = s5, t2 = s3, t3 = s1;
		p3 = t0 + t2, p4 = t1 + t3, p1 = t0 + t3, p2 = t1 + t2;
		p5 = (p3 + p4) * fix(1.175875602f);
		t0 = t0 * fix(0.298631336f);
		t1 = t1 * fix(2.053119869f);
		t2 = t2 * fix(3.072711026f);
		t3 = t3 * fix(1.501321110f);
		p1 = p5 + p1 * fix(-0.899976223f);
		p2 = p5 + p2 * fix(-2.562915447f);
		p3 = p3 * fix(-1.961570560f);
		p4 = p4 * fix(-0.390180644f);
		r.t3 = t3 + p1 + p4, r.t2 = t2 + p2 + p3, r.t1 = t1 + p2 + p4, r.t0 = t0 + p1 + p3;
		return r;
	}
	static uint8_t clamp8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }
	static void idct(uint8_t* out, int stride, const int16_t* d)
	{
		int val[64];
		for (int i = 0; i < 8; i++)
		{
			const int16_t* c = d + i;
			int* v = val + i;
			if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0)
			{
				int dc = c[0] * 4;
				for (int k = 0; k < 8; k++) v[k * 8] = dc;
				continue;
			}
			Idct1D r = idct1d(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56]);
			r.x0 += 512, r.x1 += 512, r.x2 += 512, r.x3 += 512;
			v[0] = (r.x0 + r.t3) >> 10, v[56] = (r.x0 - r.t3) >> 10;
			v[8] = (r.x1 + r.t2) >> 10, v[48] = (r.x1 - r.t2) >> 10;
			v[16] = (r.x2 + r.t1) >> 10, v[40] = (r.x2 - r.t1) >> 10;
			v[24] = (r.x3 + r.t0) >> 10, v[32] = (r.x3 - r.t0) >> 10;
		}
		for (int i = 0; i < 8; i++)
		{
			const int* v = val + i * 8;
			uint8_t* o = out + (size_t)i * stride;
			Idct1D r = idct1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
			const int bias = 65536 + (128 << 17);
			r.x0 += bias, r.x1 += bias, r.x2 += bias, r.x3 += bias;
			o[0] = clamp8((r.x0 + r.t3) >> 17), o[7] = clamp8((r.x0 - r.t3) >> 17);
			o[1] = clamp8((r.x1 + r.t2) >> 17), o[6] = clamp8((r.x1 - r.t2) >> 17);
			o[2] = clamp8((r.x2 + r.t1) >> 17), o[5] = clamp8((r.x2 - r.t1) >> 17);
			o[3] = clamp8((r.x3 + r.t0) >> 17), o[4] = clamp8((r.x3 - r.t0) >> 17);
		}
	}

	// ---- up-sampling of one output row: returns the row to read (w_lores * hs samples)
	static const uint8_t* upsampleRow(uint8_t* buf, const uint8_t* nearRow, const uint8_t* farRow, int w, int hs, int vs)
	{
		if (hs == 1 && vs == 1) return nearRow;
		if (hs == 1 && vs == 2)
		{
			for (int i = 0; i < w; i++) buf[i] = (uint8_t)((3 * nearRow[i] + farRow[i] + 2) >> 2);
			return buf;
		}
		if (hs == 2 && vs == 1)
		{
			const uint8_t* in = nearRow;
			if (w == 1)
			{
				buf[0] = buf[1] = in[0];
				return buf;
			}
			buf[0] = in[0];
			buf[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
			int i;
			for (i = 1; i < w - 1; i++)
			{
				int n = 3 * in[i] + 2;
				buf[i * 2] = (uint8_t)((n + in[i - 1]) >> 2);
				buf[i * 2 + 1] = (uint8_t)((n + in[i + 1]) >> 2);
			}
			buf[i * 2] = (uint8_t)((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
			buf[i * 2 + 1] = in[w - 1];
			return buf;
		}
		if (hs == 2 && vs == 2)
		{
			if (w == 1)
			{
				buf[0] = buf[1] = (uint8_t)((3 * nearRow[0] + farRow[0] + 2) >> 2);
				return buf;
			}
			int t1 = 3 * nearRow[0] + farRow[0];
			buf[0] = (uint8_t)((t1 + 2) >> 2);
			for (int i = 1; i < w; i++)
			{
				int t0 = t1;
				t1 = 3 * nearRow[i] + farRow[i];
				buf[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
				buf[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
			}
			buf[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
			return buf;
		}
		for (int i = 0; i < w; i++) // any other ratio: replicate horizontally, nearest row vertically
			for (int j = 0; j < hs; j++) buf[i * hs + j] = nearRow[i];
		return buf;
	}

	static int fixColour(float x) { return ((int)(x * 4096.0f + 0.5f)) << 8; }
	static uint8_t blinn(uint8_t x, uint8_t y)
	{
		unsigned t = x * y + 128;
		return (uint8_t)((t + (t >> 8)) >> 8);
	}

	bool output(int& width, int& height, int& channels, std::vector<unsigned char>& out)
	{
		width = imgW, height = imgH;
		int n = nComp >= 3 ? 3 : 1;
		channels = n;
		bool isRGB = nComp == 3 && (rgbIds == 3 || (adobeTransform == 0 && !jfif));
		int decodeN = nComp;
		struct Resample
		{
			int hs, vs, wLores, ystep, ypos;
			const uint8_t *line0, *line1;
			std::vector<uint8_t> buf;
		} rs[4];
		for (int k = 0; k < decodeN; k++)
		{
			rs[k].hs = hMax / comp[k].h, rs[k].vs = vMax / comp[k].v;
			rs[k].ystep = rs[k].vs >> 1;
			rs[k].wLores = (imgW + rs[k].hs - 1) / rs[k].hs;
			rs[k].ypos = 0;
			rs[k].line0 = rs[k].line1 = comp[k].data.data();
			rs[k].buf.assign((size_t)imgW + 8, 0);
		}
		out.assign((size_t)n * imgW * imgH, 0);
		const uint8_t* rows[4] = {nullptr, nullptr, nullptr, nullptr};
		for (int j = 0; j < imgH; j++)
		{
			uint8_t* o = &out[(size_t)n * imgW * j];
			for (int k = 0; k < decodeN; k++)
			{
				Resample& r = rs[k];
				bool bottom = r.ystep >= (r.vs >> 1);
				rows[k] = upsampleRow(r.buf.data(), bottom ? r.line1 : r.line0, bottom ? r.line0 : r.line1, r.wLores, r.hs, r.vs);
				if (++r.ystep >= r.vs)
				{
					r.ystep = 0;
					r.line0 = r.line1;
					if (++r.ypos < comp[k].y) r.line1 += comp[k].w2;
				}
			}
			if (n == 1)
			{
				for (int i = 0; i < imgW; i++) o[i] = rows[0][i];
				continue;
			}
			if (nComp == 3 && isRGB)
			{
				for (int i = 0; i < imgW; i++) o[i * 3] = rows[0][i], o[i * 3 + 1] = rows[1][i], o[i * 3 + 2] = rows[2][i];
				continue;
			}
			if (nComp == 4 && adobeTransform == 0) // CMYK
			{
				for (int i = 0; i < imgW; i++)
				{
					uint8_t m = rows[3][i];
					o[i * 3] = blinn(rows[0][i], m), o[i * 3 + 1] = blinn(rows[1][i], m), o[i * 3 + 2] = blinn(rows[2][i], m);
				}
				continue;
			}
			for (int i = 0; i < imgW; i++)
			{
				int yf = (rows[0][i] << 20) + (1 << 19);
				int cr = rows[2][i] - 128, cb = rows[1][i] - 128;
				int r = yf + cr * fixColour(1.40200f);
				int g = yf + (cr * -fixColour(0.71414f)) + ((cb * -fixColour(0.34414f)) & 0xffff0000);
				int b = yf + cb * fixColour(1.77200f);
				r >>= 20, g >>= 20, b >>= 20;
				o[i * 3] = clamp8(r), o[i * 3 + 1] = clamp8(g), o[i * 3 + 2] = clamp8(b);
			}
			if (nComp == 4 && adobeTransform == 2) // YCCK
			{
				for (int i = 0; i < imgW; i++)
				{
					uint8_t m = rows[3][i];
					o[i * 3] = blinn((uint8_t)(255 - o[i * 3]), m), o[i * 3 + 1] = blinn((uint8_t)(255 - o[i * 3 + 1]), m);
					o[i * 3 + 2] = blinn((uint8_t)(255 - o[i * 3 + 2]), m);
				}
			}
		}
		return true;
	}
};

inline bool decodeJPEG(const std::vector<unsigned char>& file, int& w, int& h, int& channels, std::vector<unsigned char>& out)
{
	JpegDecoder d;
	return d.decode(file, w, h, channels, out);
}

} // namespace rtb_img
