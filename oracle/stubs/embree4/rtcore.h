// TEST INFRASTRUCTURE (oracle/): declarations of the 16 Embree-4 C-API entry
// points and POD types that the reference headers use
// (/root/reference/include/viennaray/rayTraceKernel.hpp:34-45,91,166,425;
//  rayGeometryDisk.hpp:22-29; rayBoundary.hpp:167-242; rayTrace.hpp:17-29).
// Embree 4.3.3 itself is not vendored in the reference and not installed
// here, so the symbols are supplied by oracle/mini_rtc.cpp -- a SUBSTITUTE
// intersector, not Embree.  Field order of RTCRay/RTCHit follows the public
// Embree API documentation because the reference reinterprets the structs
// as packed float quadruples (rayUtil.hpp:218-245).
#pragma once
#include <cstddef>
#include <cstdint>

#define RTC_INVALID_GEOMETRY_ID ((unsigned int)-1)
#define RTC_MAX_INSTANCE_LEVEL_COUNT 1

extern "C" {

typedef struct MiniRtcDevice *RTCDevice;
typedef struct MiniRtcScene *RTCScene;
typedef struct MiniRtcGeometry *RTCGeometry;

enum RTCError { RTC_ERROR_NONE = 0, RTC_ERROR_UNKNOWN = 1 };
enum RTCDeviceProperty { RTC_DEVICE_PROPERTY_VERSION = 0 };
enum RTCSceneFlags { RTC_SCENE_FLAG_NONE = 0 };
enum RTCBuildQuality {
  RTC_BUILD_QUALITY_LOW = 0,
  RTC_BUILD_QUALITY_MEDIUM = 1,
  RTC_BUILD_QUALITY_HIGH = 2
};
enum RTCGeometryType {
  RTC_GEOMETRY_TYPE_TRIANGLE = 0,
  RTC_GEOMETRY_TYPE_ORIENTED_DISC_POINT = 52
};
enum RTCBufferType {
  RTC_BUFFER_TYPE_INDEX = 0,
  RTC_BUFFER_TYPE_VERTEX = 1,
  RTC_BUFFER_TYPE_NORMAL = 3
};
enum RTCFormat {
  RTC_FORMAT_UINT3 = 0x5003,
  RTC_FORMAT_FLOAT3 = 0x9003,
  RTC_FORMAT_FLOAT4 = 0x9004
};

struct alignas(16) RTCRay {
  float org_x, org_y, org_z, tnear;
  float dir_x, dir_y, dir_z, time;
  float tfar;
  unsigned int mask, id, flags;
};

struct alignas(16) RTCHit {
  float Ng_x, Ng_y, Ng_z;
  float u, v;
  unsigned int primID, geomID;
  unsigned int instID[RTC_MAX_INSTANCE_LEVEL_COUNT];
};

struct alignas(16) RTCRayHit {
  struct RTCRay ray;
  struct RTCHit hit;
};

RTCDevice rtcNewDevice(const char *config);
void rtcReleaseDevice(RTCDevice device);
long rtcGetDeviceProperty(RTCDevice device, enum RTCDeviceProperty prop);
enum RTCError rtcGetDeviceError(RTCDevice device);

RTCScene rtcNewScene(RTCDevice device);
void rtcSetSceneFlags(RTCScene scene, enum RTCSceneFlags flags);
void rtcSetSceneBuildQuality(RTCScene scene, enum RTCBuildQuality quality);
unsigned int rtcAttachGeometry(RTCScene scene, RTCGeometry geometry);
void rtcJoinCommitScene(RTCScene scene);
void rtcReleaseScene(RTCScene scene);

RTCGeometry rtcNewGeometry(RTCDevice device, enum RTCGeometryType type);
void *rtcSetNewGeometryBuffer(RTCGeometry geometry, enum RTCBufferType type,
                              unsigned int slot, enum RTCFormat format,
                              size_t byteStride, size_t itemCount);
void rtcSetGeometryMask(RTCGeometry geometry, unsigned int mask);
void rtcCommitGeometry(RTCGeometry geometry);
void rtcReleaseGeometry(RTCGeometry geometry);

void rtcIntersect1(RTCScene scene, struct RTCRayHit *rayhit);

// --- substitute-intersector extras (not Embree API), used by the oracle
// driver only: statistics of the last committed scene.
size_t miniRtcSceneNodeCount(RTCScene scene);
}
