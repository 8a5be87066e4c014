// b200_io.cpp -- dependency-free writer / reader for the reference's segment YAML schema (see io.h).
#include "mav_trajectory_generation/io.h"
#include "mav_trajectory_generation/trajectory_sampling.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <vector>

namespace mav_trajectory_generation {

std::string segmentsToYamlString(const Segment::Vector& segments) {
  std::ostringstream out;
  char buf[64];
  out << "segments:";
  if (segments.empty()) out << " []";
  out << "\n";
  for (const Segment& s : segments) {
    out << "  - N: " << s.N() << "\n";
    out << "    D: " << s.D() << "\n";
    out << "    time: " << s.getTimeNSec() << "  # [ns]\n";
    out << "    coefficients:\n";
    for (int d = 0; d < s.D(); ++d) {
      const Eigen::VectorXd c = s[d].getCoefficients();
      out << "      - [";
      for (int j = 0; j < s.N(); ++j) {
        std::snprintf(buf, sizeof(buf), "%.17g", c[j]);
        out << (j ? ", " : "") << buf;
      }
      out << "]\n";
    }
  }
  return out.str();
}

bool segmentsToFile(const std::string& filename, const Segment::Vector& segments) {
  std::ofstream fout(filename);
  if (!fout) return false;
  fout << segmentsToYamlString(segments);
  return static_cast<bool>(fout);
}

namespace {

std::string stripComment(const std::string& line) {
  const size_t hash = line.find('#');
  std::string s = hash == std::string::npos ? line : line.substr(0, hash);
  while (!s.empty() && std::isspace(static_cast<unsigned char>(s.back()))) s.pop_back();
  return s;
}

// "key: value" after optional "- "; returns false if the line is not of that form
bool keyValue(const std::string& body, std::string* key, std::string* value) {
  const size_t colon = body.find(':');
  if (colon == std::string::npos) return false;
  *key = body.substr(0, colon);
  *value = body.substr(colon + 1);
  auto trim = [](std::string* s) {
    size_t a = 0;
    while (a < s->size() && std::isspace(static_cast<unsigned char>((*s)[a]))) ++a;
    size_t b = s->size();
    while (b > a && std::isspace(static_cast<unsigned char>((*s)[b - 1]))) --b;
    *s = s->substr(a, b - a);
  };
  trim(key);
  trim(value);
  return !key->empty();
}

bool parseFlowSequence(const std::string& text, std::vector<double>* values) {
  const size_t open = text.find('['), close = text.rfind(']');
  if (open == std::string::npos || close == std::string::npos || close < open) return false;
  values->clear();
  std::string inner = text.substr(open + 1, close - open - 1);
  std::stringstream ss(inner);
  std::string item;
  while (std::getline(ss, item, ',')) {
    char* end = nullptr;
    const double v = std::strtod(item.c_str(), &end);
    if (end == item.c_str()) return false;
    values->push_back(v);
  }
  return true;
}

struct PendingSegment {
  int N = -1, D = -1;
  bool has_time = false;
  uint64_t time_ns = 0;
  std::vector<std::vector<double> > coefficients;
  bool complete() const { return N > 0 && D > 0 && has_time && static_cast<int>(coefficients.size()) == D; }
};

bool flush(const PendingSegment& p, Segment::Vector* segments) {
  if (!p.complete()) return false;  // wrong format, missing elements (reference io.cpp:213-215)
  Segment segment(p.N, p.D);
  segment.setTimeNSec(p.time_ns);
  for (int d = 0; d < p.D; ++d) {
    if (static_cast<int>(p.coefficients[d].size()) != p.N) return false;
    Eigen::VectorXd c(p.N);
    for (int j = 0; j < p.N; ++j) c[j] = p.coefficients[d][j];
    segment[d] = Polynomial(p.N, c);
  }
  segments->push_back(segment);
  return true;
}

}  // namespace

bool segmentsFromYamlString(const std::string& yaml, Segment::Vector* segments) {
  CHECK_NOTNULL(segments)->clear();
  std::istringstream in(yaml);
  std::string raw;
  bool saw_root = false, in_segment = false, in_coefficients = false;
  PendingSegment cur;
  while (std::getline(in, raw)) {
    const std::string line = stripComment(raw);
    size_t indent = 0;
    while (indent < line.size() && line[indent] == ' ') ++indent;
    if (indent == line.size()) continue;  // blank
    std::string body = line.substr(indent);
    if (!saw_root) {
      std::string key, value;
      if (!keyValue(body, &key, &value) || key != "segments") return false;  // no segments element
      saw_root = true;
      continue;
    }
    bool dash = false;
    if (body.size() >= 2 && body[0] == '-' && body[1] == ' ') {
      dash = true;
      body = body.substr(2);
      while (!body.empty() && body[0] == ' ') body = body.substr(1);
    }
    if (dash && !body.empty() && body[0] == '[') {  // one dimension's coefficients
      if (!in_segment || !in_coefficients) return false;
      std::vector<double> values;
      if (!parseFlowSequence(body, &values)) return false;
      cur.coefficients.push_back(values);
      continue;
    }
    std::string key, value;
    if (!keyValue(body, &key, &value)) return false;
    if (dash) {  // a new element of the segments sequence starts
      if (in_segment && !flush(cur, segments)) return false;
      cur = PendingSegment();
      in_segment = true;
    }
    if (!in_segment) return false;
    in_coefficients = false;
    if (key == "N") {
      cur.N = std::atoi(value.c_str());
    } else if (key == "D") {
      cur.D = std::atoi(value.c_str());
    } else if (key == "time") {
      cur.time_ns = std::strtoull(value.c_str(), nullptr, 10);
      cur.has_time = true;
    } else if (key == "coefficients") {
      in_coefficients = true;
      if (!value.empty()) return false;  // only the block form the reference emits is supported
    } else {
      return false;
    }
  }
  if (!saw_root) return false;
  if (in_segment && !flush(cur, segments)) return false;
  return true;
}

bool segmentsFromFile(const std::string& filename, Segment::Vector* segments) {
  CHECK_NOTNULL(segments);
  std::ifstream in(filename);
  if (!in.good()) return false;
  std::stringstream buffer;
  buffer << in.rdbuf();
  return segmentsFromYamlString(buffer.str(), segments);
}

bool sampledTrajectoryStatesToFile(const std::string& filename, const Trajectory& trajectory) {
  const double sampling_time = 0.01;
  mav_msgs::EigenTrajectoryPoint::Vector points;
  if (!sampleWholeTrajectory(trajectory, sampling_time, &points)) return false;
  const int dim = 3, cols = 8 * dim + 3;
  std::vector<double> table(points.size() * size_t(cols), 0.0);
  for (size_t i = 0; i < points.size(); ++i) {
    const mav_msgs::EigenTrajectoryPoint& st = points[i];
    double* row = &table[i * cols];
    row[0] = static_cast<double>(st.time_from_start_ns);
    for (int d = 0; d < dim; ++d) {
      row[1 + d] = st.position_W[d];
      row[1 + dim + d] = st.velocity_W[d];
      row[1 + 2 * dim + d] = st.acceleration_W[d];
      row[1 + 3 * dim + d] = st.jerk_W[d];
      row[1 + 4 * dim + d] = st.snap_W[d];
      row[2 + 6 * dim + d] = st.angular_velocity_W[d];
      row[2 + 7 * dim + d] = st.angular_acceleration_W[d];
    }
    row[1 + 5 * dim] = st.orientation_W_B.w;
    row[2 + 5 * dim] = st.orientation_W_B.x;
    row[3 + 5 * dim] = st.orientation_W_B.y;
    row[4 + 5 * dim] = st.orientation_W_B.z;
  }
  double accumulated = 0.0;  // end time of segment j in the last column of row j
  for (size_t j = 0; j < trajectory.segments().size() && j < points.size(); ++j) {
    accumulated += trajectory.segments()[j].getTime();
    table[j * cols + (2 + 8 * dim)] = accumulated;
  }
  // text layout of an Eigen matrix streamed with the default format: %g-style 6 significant digits, one space between
  // coefficients, every coefficient right-aligned to the widest one of the whole matrix
  std::vector<std::string> cells(table.size());
  size_t width = 0;
  for (size_t k = 0; k < table.size(); ++k) {
    std::ostringstream cell;
    cell << table[k];
    cells[k] = cell.str();
    width = std::max(width, cells[k].size());
  }
  std::ofstream fs(filename);
  if (!fs) return false;
  for (size_t i = 0; i < points.size(); ++i) {
    for (int c = 0; c < cols; ++c) {
      const std::string& cell = cells[i * cols + c];
      if (c) fs << ' ';
      fs << std::string(width - cell.size(), ' ') << cell;
    }
    if (i + 1 < points.size()) fs << '\n';
  }
  return static_cast<bool>(fs);
}

}  // namespace mav_trajectory_generation
