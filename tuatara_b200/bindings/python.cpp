// pytuatara: the reference's Python module (bindings/python.cpp:1-58) re-created on the B200 path.
// Same module name, function name, keyword names and return shape:
//   pytuatara.image_to_data(image, weights_dir, outputs_dir) -> list[dict(text=str, bbox=list[float])]
// The reference copies the buffer into a cv::Mat (buffer_to_mat, :12-28); here the numpy buffer is
// passed straight to the C ABI as an ImageView.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "tuatara.h"

namespace py = pybind11;

static tuatara::ImageView buffer_to_view(const py::array_t<unsigned char, py::array::c_style | py::array::forcecast>& input) {
  py::buffer_info buf = input.request();
  if (buf.ndim != 3) throw std::runtime_error("Input array should have 3 dimensions");  // python.cpp:15-17
  tuatara::ImageView v;
  v.data = static_cast<const uint8_t*>(buf.ptr);
  v.rows = static_cast<int>(buf.shape[0]);
  v.cols = static_cast<int>(buf.shape[1]);
  v.channels = static_cast<int>(buf.shape[2]);
  v.step = static_cast<size_t>(buf.strides[0]);
  if (v.channels != 3) throw std::runtime_error("Input array should have 3 channels");
  return v;
}

static py::dict output_item_to_dict(const OutputItem& item) {  // python.cpp:36-41
  py::dict d;
  d["text"] = item.text;
  d["bbox"] = item.bbox;
  return d;
}

static py::list image_to_data_wrapper(py::array_t<unsigned char, py::array::c_style | py::array::forcecast> image_data,
                                      std::string weights_dir, std::string output_dir) {
  const tuatara::ImageView img = buffer_to_view(image_data);
  std::vector<OutputItem> items;
  {
    py::gil_scoped_release release;
    items = image_to_data(img, weights_dir, output_dir);
  }
  py::list result;
  for (const auto& item : items) result.append(output_item_to_dict(item));
  return result;
}

static py::list images_to_data_wrapper(std::vector<py::array_t<unsigned char, py::array::c_style | py::array::forcecast>> images,
                                       std::string weights_dir, std::string output_dir) {
  std::vector<tuatara::ImageView> views;
  for (auto& a : images) views.push_back(buffer_to_view(a));
  std::vector<std::vector<OutputItem>> pages;
  {
    py::gil_scoped_release release;
    pages = tuatara::image_to_data_batch(views, weights_dir, output_dir);
  }
  py::list result;
  for (const auto& page : pages) {
    py::list items;
    for (const auto& item : page) items.append(output_item_to_dict(item));
    result.append(items);
  }
  return result;
}

PYBIND11_MODULE(pytuatara, m) {
  m.doc() = "Tuatara ocr";
  m.def("image_to_data", &image_to_data_wrapper, py::arg("image"), py::arg("weights_dir"), py::arg("outputs_dir"),
        "Extract text and bounding boxes from an image");
  m.def("images_to_data", &images_to_data_wrapper, py::arg("images"), py::arg("weights_dir"), py::arg("outputs_dir"),
        "Batch variant: list of images -> list of per-page results");
}
