// TEST INFRASTRUCTURE (oracle/): no-op stand-in for Intel OIDN so the reference's
// Renderer.h parses.  RayTracer::denoise (RTBase/Renderer.h:752-793) is commented out at
// every call site (:230, :850), so none of this ever runs.
#pragma once
#include <cstddef>
#include <vector>
namespace oidn
{
enum class Format { Float3 };
struct BufferRef
{
	std::vector<char> bytes;
	void* getData() { return bytes.data(); }
};
struct FilterRef
{
	void setImage(const char*, BufferRef&, Format, size_t, size_t) {}
	template <class T> void set(const char*, T) {}
	void commit() {}
	void execute() {}
};
struct DeviceRef
{
	void commit() {}
	BufferRef newBuffer(size_t n)
	{
		BufferRef b;
		b.bytes.resize(n);
		return b;
	}
	FilterRef newFilter(const char*) { return FilterRef(); }
	bool getError(const char*&) { return false; }
};
inline DeviceRef newDevice() { return DeviceRef(); }
} // namespace oidn
