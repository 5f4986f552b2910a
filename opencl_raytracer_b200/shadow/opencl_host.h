/*
 * opencl_host.h -- SHADOW of the reference's include/opencl_host.h.
 *
 * Put this directory before the reference's include/ on the include path and
 * the reference's src/render.cc compiles UNMODIFIED against the B200 path:
 * it declares `class OpenCLHost` with exactly the surface render.cc uses
 * (reference include/opencl_host.h:127-131) and forwards every call to the
 * C ABI of include/rtx_b200.h.  No OpenCL header, ICD or device is needed.
 *
 *   OpenCLHost::printInfo()            -> rtx_device_info, printed through the
 *                                         reference's own Info tree (info.h)
 *   OpenCLHost(const RayTracer &rt)    -> rtx_create   (options from rt.options,
 *                                         rt.totalWidth / rt.totalHeight)
 *   upload(faces,nodes,aabbs,verts,vn) -> rtx_upload   (blocking: render.cc:96-103
 *                                         clears the vectors right after)
 *   bool operator()()                  -> rtx_render   (launch + wait)
 *   download(float *image)             -> rtx_download
 *
 * Error behaviour restored from the reference: a failing device call prints
 * the message and exits (opencl_host.h:21-26 `check`), no device throws
 * std::runtime_error("No device found") (opencl_host.cc:30-31), operator()
 * returns false on failure so Info::measure prints "failed!" and exits
 * (info.cc:26-40).
 *
 * Build line: INTEGRATION.md.
 */
#pragma once
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "color.h"
#include "info.h"
#include "ray_tracer.h"
#include "rtx_b200.h"
#include "vec3.h"

class OpenCLHost {
	public:
		static void check(const rtx_ctx *ctx, int err) {
			if (err != RTX_OK) {
				std::cerr << "CUDA host error: " << rtx_last_error(ctx) << std::endl;
				std::exit(EXIT_FAILURE);
			}
		}
		OpenCLHost(const RayTracer &rt) : rt(rt), ctx(nullptr) {
			static_assert(sizeof(Vec3f) == 16, "Vec3f must keep the float4 layout the device arrays use");
			rtx_options o{};
			o.width = rt.options.width;
			o.height = rt.options.height;
			o.focal_length = rt.options.focalLength;
			o.n_super_samples = rt.options.nSuperSamples;
			o.enable_shading = rt.options.enableShading;
			o.enable_ao = rt.options.enableAO;
			o.ao_max_distance = rt.options.aoMaxDistance;
			o.ao_num_samples = rt.options.aoNumSamples;
			o.ao_method = (int32_t) rt.options.aoMethod;
			o.ao_alpha_min = rt.options.aoAlphaMin;
			o.ao_alpha_max = rt.options.aoAlphaMax;
			o.bvh_method = (int32_t) rt.options.bvhMethod;
			o.total_width = rt.totalWidth;
			o.total_height = rt.totalHeight;
			const char *dev = std::getenv("RTX_DEVICE");
			o.device = dev ? std::atoi(dev) : 0;
			const int err = rtx_create(&ctx, &o);
			if (err == RTX_ERR_NO_DEVICE)
				throw std::runtime_error("No device found");
			check(nullptr, err);
			rtx_device_info_t info;
			if (rtx_device_info(o.device, &info) == RTX_OK)
				std::cout << Color::WHITE << "Using Device \"" << info.name << "\"." << Color::RESET << std::endl << std::endl;
		}
		~OpenCLHost() {
			rtx_destroy(ctx);
		}
		OpenCLHost(const OpenCLHost &) = delete;
		OpenCLHost &operator=(const OpenCLHost &) = delete;
		void upload(const std::vector<uint32_t> &faces, const std::vector<uint32_t> &nodes, const std::vector<Vec3f> &aabbs, const std::vector<Vec3f> &vertices, const std::vector<Vec3f> &vnormals) {
			const std::size_t mem = faces.size() * sizeof(uint32_t) + nodes.size() * sizeof(uint32_t)
				+ (aabbs.size() + vertices.size() + vnormals.size()) * sizeof(Vec3f)
				+ (std::size_t) rt.totalWidth * rt.totalHeight * sizeof(float);
			std::cout << "Requested " << mem / 1024 << " kB of memory." << std::endl;
			check(ctx, rtx_upload(ctx,
				faces.data(), faces.size(),
				nodes.data(), nodes.size(),
				reinterpret_cast<const float *>(aabbs.data()), aabbs.size(),
				reinterpret_cast<const float *>(vertices.data()), vertices.size(),
				reinterpret_cast<const float *>(vnormals.data()), vnormals.size()));
		}
		bool operator()() {
			const int err = rtx_render(ctx);
			if (err != RTX_OK)
				std::cerr << "CUDA host error: " << rtx_last_error(ctx) << std::endl;
			return err == RTX_OK;
		}
		void download(float *image) {
			check(ctx, rtx_download(ctx, image));
		}
		static void printInfo() {
			Info info;
			info.setTitle("Hardware information");
			int n = 0;
			rtx_device_count(&n);
			Info platformInfo;
			platformInfo.setTitle("Platform #0");
			platformInfo.add("Name", std::string("NVIDIA CUDA (rtx_b200)"));
			platformInfo.add("Devices", n);
			for (int j = 0; j < n; ++j) {
				rtx_device_info_t d;
				if (rtx_device_info(j, &d) != RTX_OK)
					continue;
				Info deviceInfo;
				std::stringstream deviceTitle;
				deviceTitle << "Device #" << j;
				deviceInfo.setTitle(deviceTitle.str());
				deviceInfo.add("Name", std::string(d.name));
				std::stringstream cc;
				cc << d.cc_major << "." << d.cc_minor;
				deviceInfo.add("Compute capability", cc.str());
				deviceInfo.add("Driver version", d.driver_version);
				deviceInfo.add("Runtime version", d.runtime_version);
				deviceInfo.add("Type", std::string("GPU"));
				deviceInfo.add("Max compute units", d.sm_count);
				deviceInfo.add("Global memory size (MiB)", d.global_mem_bytes >> 20);
				deviceInfo.add("L2 cache size (MiB)", d.l2_bytes >> 20);
				deviceInfo.add("Local memory size (B)", d.smem_per_block_optin_bytes);
				platformInfo.add(deviceInfo);
			}
			info.add(platformInfo);
			std::cout << std::endl;
			std::cout << info.str();
		}
	private:
		const RayTracer &rt;
		rtx_ctx *ctx;
};
