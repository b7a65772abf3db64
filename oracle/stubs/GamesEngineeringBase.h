// TEST INFRASTRUCTURE (oracle/): headless stand-in for the reference's Win32/D3D11
// platform header so that the UNMODIFIED reference headers under /root/reference/RTBase
// compile with g++ on Linux.  Only what the path-tracing translation unit touches is
// provided (reference uses: RTBase/Renderer.h:35,45,52-53,77,871,897; RTBase/Imaging.h:7).
// Nothing here is product code; the product never includes this file.
#pragma once
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#ifndef sprintf_s
#define sprintf_s snprintf
#endif
#define VK_ESCAPE 0x1B

// The reference relies on the <windows.h> min/max macros with mixed int/unsigned
// arguments (Renderer.h:353, 798-799).  Same semantics: compare after the usual
// arithmetic conversions, return the common type.  min(a,b) = a<b ? a : b.
template <class A, class B>
inline std::common_type_t<A, B> min(A a, B b)
{
	typedef std::common_type_t<A, B> C;
	return ((C)a < (C)b) ? (C)a : (C)b;
}
template <class A, class B>
inline std::common_type_t<A, B> max(A a, B b)
{
	typedef std::common_type_t<A, B> C;
	return ((C)a > (C)b) ? (C)a : (C)b;
}

struct SYSTEM_INFO
{
	unsigned int dwNumberOfProcessors;
};
// Worker-thread count for RayTracer::init.  0 = use every hardware thread.
inline int& rtb_ref_thread_override()
{
	static int n = 0;
	return n;
}
inline void GetSystemInfo(SYSTEM_INFO* info)
{
	int n = rtb_ref_thread_override();
	if (n <= 0)
	{
		const char* e = getenv("RTB_REF_THREADS");
		n = e ? atoi(e) : 0;
	}
	if (n <= 0) n = (int)std::thread::hardware_concurrency();
	if (n <= 0) n = 1;
	info->dwNumberOfProcessors = (unsigned int)n;
}

namespace GamesEngineeringBase
{
class Window
{
	std::vector<unsigned char> back;
	unsigned int w = 0, h = 0;

public:
	void create(unsigned int _w, unsigned int _h, const std::string&, float = 1.0f)
	{
		w = _w;
		h = _h;
		back.assign((size_t)w * h * 3, 0);
	}
	void checkInput() {}
	void clear() {}
	void present() {}
	bool keyPressed(int) const { return false; }
	void draw(unsigned int x, unsigned int y, unsigned char r, unsigned char g, unsigned char b)
	{
		size_t i = ((size_t)y * w + x) * 3;
		back[i] = r;
		back[i + 1] = g;
		back[i + 2] = b;
	}
	unsigned int getWidth() const { return w; }
	unsigned int getHeight() const { return h; }
	unsigned char* getBackBuffer() { return back.data(); }
};
class Timer
{
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();

public:
	void reset() { t0 = std::chrono::steady_clock::now(); }
	float dt()
	{
		auto t1 = std::chrono::steady_clock::now();
		float s = std::chrono::duration<float>(t1 - t0).count();
		t0 = t1;
		return s;
	}
};
} // namespace GamesEngineeringBase
